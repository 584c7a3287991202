#!/bin/bash
# round-2 call 2: GN fused finalize tests + A/B timings of one UNet call at P = 64
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_parity_s50_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
for ff in 0 1; do for rv in 0 1; do
  echo "== WDM_GN_FUSED_FINALIZE=$ff WDM_GN_REVERSE=$rv" >> gpurun_out/c2_ab.txt
  WDM_GN_FUSED_FINALIZE=$ff WDM_GN_REVERSE=$rv timeout 200 python tools/profile_unet.py --patches 64 --iters 30 --time 2>&1 | grep "ms/forward" >> gpurun_out/c2_ab.txt
done; done
tail -4 gpurun_out/c2_pytest.log; cat gpurun_out/c2_ab.txt
