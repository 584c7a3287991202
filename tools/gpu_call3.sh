#!/bin/bash
# round-2 call 3: full GPU suite, the default bench line, then the evidence pass on this binary
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 600 python bench.py > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?" >> gpurun_out/c3_bench.err
bash tools/evidence.sh r02a > gpurun_out/c3_evidence.log 2>&1
tail -3 gpurun_out/c3_pytest.log; tail -c 400 gpurun_out/c3_bench.json; tail -5 gpurun_out/c3_evidence.log
