#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hfrm_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
timeout 300 python tools/bench_hfrm.py --precisions bf16 > gpurun_out/c17_hfrm.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"hfrm_|gemm_tc" -c 700 --csv \
    --log-file gpurun_out/c17_hfrm_launches.csv python tools/bench_hfrm.py --precisions bf16 --iters 1 > gpurun_out/c17_ncu.log 2>&1
tail -3 gpurun_out/c17_pytest.log; cat gpurun_out/c17_hfrm.txt
