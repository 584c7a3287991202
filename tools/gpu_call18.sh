#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"hfrm_dw_gate" -s 24 -c 1 -f -o gpurun_out/c18_dw python tools/bench_hfrm.py --precisions bf16 --iters 1 > gpurun_out/c18_ncu.log 2>&1
ncu -i gpurun_out/c18_dw.ncu-rep --page raw --csv > gpurun_out/c18_dw_raw.csv 2>&1
ncu -i gpurun_out/c18_dw.ncu-rep --page details > gpurun_out/c18_dw_details.txt 2>&1
rm -f gpurun_out/c18_dw.ncu-rep
grep -c "" gpurun_out/c18_dw_raw.csv
