#!/bin/bash
# round-2 call 15: default bench line (HFRM on), HFRM bypassed for comparison with round 1, config 5
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err; echo "rc=$?" >> gpurun_out/c15_bench.err
timeout 600 python bench.py --bypass-hfrm --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/c15_bench_bypass.json 2> gpurun_out/c15_bench_bypass.err
timeout 600 python bench.py --config 5 --steps 1 --warmup 1 --no-parity --no-gpu-baseline --no-cpu-baseline > gpurun_out/c15_bench_cfg5.json 2> gpurun_out/c15_bench_cfg5.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/c15_bench_ref.json 2> gpurun_out/c15_bench_ref.err
tail -2 gpurun_out/c15_bench.err; head -c 300 gpurun_out/c15_bench.json; echo; head -c 300 gpurun_out/c15_bench_bypass.json; echo; head -c 300 gpurun_out/c15_bench_cfg5.json; echo; head -c 600 gpurun_out/c15_bench_ref.json
