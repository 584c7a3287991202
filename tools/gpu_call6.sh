#!/bin/bash
# round-2 call 6: HFRM engine tests + timing + per-kernel launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hfrm_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
timeout 300 python tools/bench_hfrm.py > gpurun_out/c6_hfrm.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:hfrm_ -c 400 --csv \
    --log-file gpurun_out/c6_hfrm_launches.csv python tools/bench_hfrm.py --precisions bf16 --iters 1 > gpurun_out/c6_ncu.log 2>&1
tail -4 gpurun_out/c6_pytest.log; cat gpurun_out/c6_hfrm.txt
