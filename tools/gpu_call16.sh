#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/graph_ab.py > gpurun_out/c16_graph_ab.txt 2>&1
for mp in 148 160 148 160; do
  timeout 300 python bench.py --config 5 --steps 1 --warmup 1 --max-patches $mp --no-parity --no-gpu-baseline --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 max_patches $mp', round(b['value'],3), 'img/s', b['clocks']['sm_mhz'], 'MHz frac', round(b['roofline']['frac'],3))" >> gpurun_out/c16_cfg5_ab.txt
done
cat gpurun_out/c16_graph_ab.txt | grep GRAPH; cat gpurun_out/c16_cfg5_ab.txt
