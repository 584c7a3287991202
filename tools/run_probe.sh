#!/bin/bash
# GPU-box driver (run through gpurun): the round-end sequence the driver runs -> gpurun_out/
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/run_probe.sh'
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) >> $O
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1) >> $O
timeout 200 python tools/profile_unet.py --patches 64 --iters 10 --time 2>&1 | grep "^P=" >> $O
cat $O
