#!/bin/bash
# GPU-box driver (run through gpurun): the round-end sequence the driver runs -> gpurun_out/
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/run_probe.sh'
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) >> $O
timeout 300 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_final2.json 2>> $O
tail -2 $O
