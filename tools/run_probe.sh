#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) >> $O
timeout 200 python tools/profile_unet.py --patches 64 --iters 10 --time --spans 2>&1 | grep -E "^P=| 16  12 | 16  24 |   8  12 |   8  24 " >> $O
cat $O
