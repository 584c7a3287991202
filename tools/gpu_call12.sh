#!/bin/bash
# round-2 call 12: evidence pass (launch list, DRAM bytes, ncu --set full summaries, DWT DRAM bytes) on the current binary
bash tools/evidence.sh r02b > gpurun_out/c12_evidence.log 2>&1
tail -30 gpurun_out/c12_evidence.log
