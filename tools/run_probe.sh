#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(WDM_TC_PAIRBAR=1 timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -k "tc or unet" 2>&1 | tail -3) >> $O
for v in 1 0; do
WDM_TC_PAIRBAR=$v timeout 100 python tools/tc_probe.py 2>&1 | grep "256->256\|512->512 @16x16 taps=9 full=0\|768->768" >> $O
WDM_TC_PAIRBAR=$v timeout 200 python tools/profile_unet.py --patches 64 --iters 10 --time 2>&1 | grep "^P=" >> $O
done
cat $O
