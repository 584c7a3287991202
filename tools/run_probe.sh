#!/bin/bash
# GPU-box probe driver (run through gpurun): prints to gpurun_out/probe.log
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -15) >> $O
timeout 100 python tools/tc_probe.py >> $O 2>&1
timeout 200 python tools/profile_unet.py --patches 64 --iters 3 --time --spans >> $O 2>&1
tail -5 $O
