#!/bin/bash
# GPU-box probe driver (run through gpurun): prints to gpurun_out/probe.log
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) >> $O
timeout 100 python tools/tc_probe.py small >> $O 2>&1
timeout 200 python tools/profile_unet.py --patches 64 --iters 3 --time >> $O 2>&1
timeout 300 python bench.py --no-cpu-baseline --steps 5 >> $O 2>&1
tail -5 $O
