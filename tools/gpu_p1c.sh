mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -p no:cacheprovider -k "splitk" > gpurun_out/p1c_pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/p1c_pytest_a.log
tail -4 gpurun_out/p1c_pytest_a.log
if grep -q "rc=0" gpurun_out/p1c_pytest_a.log; then
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_parity_s50_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/p1c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/p1c_pytest.log
tail -4 gpurun_out/p1c_pytest.log
for v in 1 0; do
WDM_TC_SPLITK=$v timeout 300 python tools/latency_small.py 2>&1 | grep "graph=" > gpurun_out/p1c_lat_$v.txt
done
cat gpurun_out/p1c_lat_1.txt; echo; cat gpurun_out/p1c_lat_0.txt
timeout 300 python tools/profile_unet.py --patches 1 --iters 5 --time --spans > gpurun_out/p1c_spans.txt 2>&1
head -14 gpurun_out/p1c_spans.txt
fi
