#!/bin/bash
# round-2 call 10: metrics kernel + restore() flow tests; ncu --set full of the two HBM-bound HFRM kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_metrics_gpu.py tests/test_sampler_gpu.py tests/test_compat_eval_flow_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
timeout 600 ncu --set full --clock-control none -k regex:"hfrm_dw_gate|hfrm_pw_strip" -s 104 -c 4 -f -o gpurun_out/c10_hfrm_full \
    python tools/bench_hfrm.py --precisions bf16 --iters 1 > gpurun_out/c10_ncu.log 2>&1
ncu -i gpurun_out/c10_hfrm_full.ncu-rep --page details > gpurun_out/c10_hfrm_details.txt 2>&1
python tools/ncu_summary.py gpurun_out/c10_hfrm_full.ncu-rep > gpurun_out/c10_hfrm_full.csv 2>&1
rm -f gpurun_out/c10_hfrm_full.ncu-rep
tail -5 gpurun_out/c10_pytest.log
