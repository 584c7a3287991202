#!/bin/bash
# GPU-box driver (run through gpurun): full validation + bench + ncu evidence -> gpurun_out/
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) >> $O
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2) >> $O
timeout 400 python bench.py --steps 5 > gpurun_out/bench_v12.json 2>> $O
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v12.json 2>> $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_v12_launches_unet_p64.csv python tools/profile_unet.py --patches 64 --iters 1 >> $O 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gemm_tc|gn_apply" -s 254 -c 16 -o gpurun_out/r01_v12_full python tools/profile_unet.py --patches 64 --iters 1 >> $O 2>&1
timeout 200 python tools/profile_unet.py --patches 64 --iters 5 --time --spans > gpurun_out/r01_v12_spans.txt 2>&1
tail -3 $O
