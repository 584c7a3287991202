#!/bin/bash
# round-2 final: full GPU suite, every bench line, then the evidence pass on the same binary
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "rc=$?" >> gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 600 python bench.py --bypass-hfrm --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/f_bench_bypass.json 2> /dev/null
timeout 600 python bench.py --config 5 --steps 1 --warmup 1 --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/f_bench_cfg5.json 2> /dev/null
timeout 600 python bench.py --precision fp32 --batch 16 --steps 2 --warmup 1 --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/f_bench_fp32.json 2> /dev/null
timeout 600 python bench.py --precision fp32_ffma --batch 16 --steps 1 --warmup 1 --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/f_bench_fp32_ffma.json 2> /dev/null
timeout 600 python bench.py --wavelet-in-unet --steps 2 --warmup 1 > gpurun_out/f_bench_wiu.json 2> /dev/null
bash tools/evidence.sh r02 > gpurun_out/f_evidence.log 2>&1
tail -3 gpurun_out/f_pytest.log; for f in f_bench f_bench_bypass f_bench_cfg5 f_bench_fp32 f_bench_fp32_ffma f_bench_wiu f_bench_ref; do head -c 160 gpurun_out/$f.json | tail -c 120; echo; done; tail -3 gpurun_out/f_evidence.log
