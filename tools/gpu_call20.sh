#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out; TAG=r02
: > $O/${TAG}_dwt_dram.csv
for k in dwt4x4_direct iwt4x4_direct; do
    timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:$k \
        -s 50 -c 8 --csv --log-file $O/${TAG}_dwt_dram_$k.csv python tools/bench_dwt.py > $O/${TAG}_dwt_ncu.log 2>&1
    grep -v "^==" $O/${TAG}_dwt_dram_$k.csv >> $O/${TAG}_dwt_dram.csv
    rm -f $O/${TAG}_dwt_dram_$k.csv
done
timeout 900 python -m pytest tests/test_sampler_gpu.py -m gpu -q -x -s -k "hfrm_branch" -p no:cacheprovider 2>&1 | tail -5
