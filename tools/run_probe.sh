#!/bin/bash
# GPU-box driver (run through gpurun): ncu evidence for the committed state -> gpurun_out/ (keep it < 64 MiB)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_v13_dram_unet_p64.csv python tools/profile_unet.py --patches 64 --iters 1 >> $O 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"gemm_tc|gn_apply" -s 133 -c 20 -o gpurun_out/r01_v13_full_a python tools/profile_unet.py --patches 64 --iters 1 >> $O 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"gemm_tc" -s 120 -c 14 -o gpurun_out/r01_v13_full_b python tools/profile_unet.py --patches 64 --iters 1 >> $O 2>&1
tail -2 $O
ls -la gpurun_out/ | tail -5
