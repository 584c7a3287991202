#!/bin/bash
# GPU-box driver (run through gpurun): the round-end sequence the driver runs, plus the evidence files -> gpurun_out/
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/run_probe.sh'
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) >> $O
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1) >> $O
timeout 400 python bench.py > gpurun_out/bench_final.json 2>> $O
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>> $O
timeout 200 python tools/profile_unet.py --patches 64 --iters 5 --time --spans > gpurun_out/spans_final.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv python tools/profile_unet.py --patches 64 --iters 1 >> $O 2>&1
tail -3 $O
