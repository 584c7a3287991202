#!/bin/bash
# round-2 call 14: tc32 (fp32 parity mode on the tensor cores): unit cases, UNet / sampler / 50-step parity suites, fp32 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -s -k "tc32 or fp32" -p no:cacheprovider > gpurun_out/c14_pytest_tc32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14_pytest_tc32.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
timeout 400 python bench.py --precision fp32 --batch 16 --steps 1 --warmup 1 --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/c14_bench_fp32.json 2> gpurun_out/c14_bench_fp32.err; echo "rc=$?" >> gpurun_out/c14_bench_fp32.err
timeout 400 python bench.py --precision fp32_ffma --batch 16 --steps 1 --warmup 1 --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/c14_bench_fp32_ffma.json 2> gpurun_out/c14_bench_fp32_ffma.err
grep -E "tc32|passed|failed|rc=|Error" gpurun_out/c14_pytest_tc32.log | head -20; tail -5 gpurun_out/c14_pytest.log; head -c 400 gpurun_out/c14_bench_fp32.json; echo; head -c 400 gpurun_out/c14_bench_fp32_ffma.json; cat gpurun_out/parity_s50.json | head -30
