#!/bin/bash
mkdir -p gpurun_out
WDM_TC_HELPER=1 timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for h in 0 1 0 1; do
  echo "== WDM_TC_HELPER=$h"
  WDM_TC_HELPER=$h timeout 200 python tools/tc_probe.py 2>&1 | grep -E "^P=64 C=(256->256 @32|512->512 @16x16 taps=9 full=1|768)"
  WDM_TC_HELPER=$h timeout 200 python tools/profile_unet.py --patches 64 --iters 30 --time 2>&1 | grep "ms/forward"
done
