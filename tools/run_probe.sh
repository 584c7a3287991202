#!/bin/bash
# GPU-box driver (run through gpurun): the round-end sequence the driver runs, plus the event span table -> gpurun_out/
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) >> $O
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1) >> $O
timeout 400 python bench.py > gpurun_out/bench_v14.json 2>> $O
timeout 200 python tools/profile_unet.py --patches 64 --iters 5 --time --spans > gpurun_out/r01_v14_spans.txt 2>&1
tail -3 $O
