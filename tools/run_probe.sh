#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
for dbg in 0 8 6 7 5; do
WDM_TC_DBG=$dbg timeout 100 python tools/tc_probe.py 2>&1 | grep "128->128.*full=1\|512->512.*taps=1 full=1" >> $O
done
cat $O
