#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -8) >> $O
timeout 100 python tools/tc_probe.py 2>&1 | grep "128->128" >> $O
WDM_TC_HALO=0 timeout 100 python tools/tc_probe.py 2>&1 | grep "128->128" >> $O
WDM_TC_DBG=9 timeout 100 python tools/tc_probe.py 2>&1 | grep "128->128" >> $O
WDM_TC_DBG=1 timeout 100 python tools/tc_probe.py 2>&1 | grep "128->128" >> $O
timeout 200 python tools/profile_unet.py --patches 64 --iters 5 --time 2>&1 | grep -v "^profile" >> $O
cat $O
