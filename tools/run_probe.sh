#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) >> $O
timeout 300 python tools/latency_small.py >> $O 2>&1
timeout 200 python tools/profile_unet.py --patches 64 --iters 10 --time 2>&1 | grep -v "^profile" >> $O
cat $O
