#!/bin/bash
# round-2 call 1: full GPU suite + the bench lines (default, config 5, fp32)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?" >> gpurun_out/c1_bench.err
timeout 400 python bench.py --config 5 --steps 1 --warmup 1 --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/c1_bench_cfg5.json 2> gpurun_out/c1_bench_cfg5.err; echo "rc=$?" >> gpurun_out/c1_bench_cfg5.err
timeout 400 python bench.py --precision fp32 --batch 16 --steps 1 --warmup 1 --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/c1_bench_fp32.json 2> gpurun_out/c1_bench_fp32.err; echo "rc=$?" >> gpurun_out/c1_bench_fp32.err
tail -5 gpurun_out/c1_pytest.log; tail -c 600 gpurun_out/c1_bench.json
