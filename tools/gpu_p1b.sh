mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/p1b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/p1b_pytest.log
for v in 1 0; do
WDM_TC_SMALL_BN=$v timeout 300 python tools/latency_small.py 2>&1 | grep "graph=" > gpurun_out/p1b_lat_$v.txt
done
timeout 300 python tools/profile_unet.py --patches 1 --iters 5 --time --spans > gpurun_out/p1b_spans.txt 2>&1
tail -3 gpurun_out/p1b_pytest.log; cat gpurun_out/p1b_lat_1.txt; echo; cat gpurun_out/p1b_lat_0.txt; head -16 gpurun_out/p1b_spans.txt
