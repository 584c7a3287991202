mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_optim_gpu.py tests/test_parity_s50_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
timeout 600 python tools/bench_train_step.py --json gpurun_out/h_train_step.json > gpurun_out/h_train_step.txt 2>&1
grep -v "^$" gpurun_out/h_pytest.log | tail -15; tail -5 gpurun_out/h_train_step.txt
