#!/bin/bash
# round-2 call 4: HFRM engine tests + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hfrm_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 300 python tools/bench_hfrm.py > gpurun_out/c4_hfrm.txt 2>&1
tail -30 gpurun_out/c4_pytest.log; cat gpurun_out/c4_hfrm.txt
