#!/bin/bash
# round-2 call 9: HFRM timing + profile, full GPU suite, default bench line with the HFRM on
mkdir -p gpurun_out
timeout 300 python tools/bench_hfrm.py > gpurun_out/c9_hfrm.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"hfrm_|gemm_tc" -c 700 --csv \
    --log-file gpurun_out/c9_hfrm_launches.csv python tools/bench_hfrm.py --precisions bf16 --iters 1 > gpurun_out/c9_ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
timeout 600 python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; echo "bench rc=$?" >> gpurun_out/c9_bench.err
cat gpurun_out/c9_hfrm.txt; tail -5 gpurun_out/c9_pytest.log; tail -3 gpurun_out/c9_bench.err; head -c 1500 gpurun_out/c9_bench.json
