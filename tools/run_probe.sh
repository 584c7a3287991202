#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_unet_gpu.py -m gpu -x -q -k "fused_softmax or groupnorm_sidecar" > gpurun_out/racecheck.log 2>&1
grep -c "" gpurun_out/racecheck.log
