#!/bin/bash
for d in 0 1 2 5; do WDM_TC_DBG=$d timeout 200 python tools/tc_probe.py 2>&1 | grep -E "^P=64 C=(256->256 @32|512->512 @16x16 taps=9 full=0|128->128 @64x64 taps=9 full=0|768)" ; done
