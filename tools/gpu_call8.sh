#!/bin/bash
# round-2 call 8: HFRM engine tests on every kernel path + timing + per-kernel launch list
mkdir -p gpurun_out
for env in "" "WDM_HFRM_RING=0 WDM_HFRM_TC=0"; do
  echo "== $env" >> gpurun_out/c8_pytest.log
  env $env timeout 600 python -m pytest tests/test_hfrm_gpu.py -m gpu -q -x -p no:cacheprovider >> gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
  echo "== $env" >> gpurun_out/c8_hfrm.txt
  env $env timeout 300 python tools/bench_hfrm.py --precisions bf16 >> gpurun_out/c8_hfrm.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"hfrm_|gemm_tc" -c 700 --csv \
    --log-file gpurun_out/c8_hfrm_launches.csv python tools/bench_hfrm.py --precisions bf16 --iters 1 > gpurun_out/c8_ncu.log 2>&1
grep -E "passed|failed|rc=|==" gpurun_out/c8_pytest.log; cat gpurun_out/c8_hfrm.txt
