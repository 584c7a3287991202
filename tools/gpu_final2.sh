# after tools/collect_evidence.py: the default bench line again, now that profiles/r02_traffic.json carries this build's id
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "rc=$?" >> gpurun_out/t_bench.err
timeout 300 python tools/sampler_timeline.py 2>&1 | tail -6 > gpurun_out/r02_sampler_timeline.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/t_smoke.log
tail -2 gpurun_out/t_smoke.log; cat gpurun_out/r02_sampler_timeline.txt; head -c 300 gpurun_out/t_bench.json
