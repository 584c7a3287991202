#!/bin/bash
WDM_TC_TRACE=1 timeout 200 python tools/tc_probe.py small 2>&1 | grep -E "tc_trace|^P=" | head -40
