#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hfrm_gpu.py tests/test_sampler_gpu.py tests/test_compat_eval_flow_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
timeout 300 python tools/bench_hfrm.py --precisions bf16 > gpurun_out/c11_hfrm.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"hfrm_dw_gate" -s 31 -c 2 -f -o gpurun_out/c11_dw_full \
    python tools/bench_hfrm.py --precisions bf16 --iters 1 > gpurun_out/c11_ncu.log 2>&1
ncu -i gpurun_out/c11_dw_full.ncu-rep --page details > gpurun_out/c11_dw_details.txt 2>&1
rm -f gpurun_out/c11_dw_full.ncu-rep
tail -3 gpurun_out/c11_pytest.log; cat gpurun_out/c11_hfrm.txt
