#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 600 python -m pytest tests -m gpu -x -q -k "wavelet or dwt or iwt or wiu" 2>&1 | tail -5) >> $O
timeout 300 python tools/bench_dwt.py 2>&1 | grep "gather\|nhwc" >> $O
cat $O
