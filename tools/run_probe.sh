#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
for v in 1 0; do
WDM_TC_PAIR192=$v timeout 100 python tools/tc_probe.py 2>&1 | grep "768->768" >> $O
WDM_TC_PAIR192=$v timeout 200 python tools/profile_unet.py --patches 64 --iters 10 --time 2>&1 | grep "^P=" >> $O
done
cat $O
