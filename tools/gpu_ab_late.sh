mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
for i in 1 2; do
for late in 1 0; do
WDM_PDL_LATE=$late timeout 600 python bench.py --bypass-hfrm --no-parity --no-gpu-baseline --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/g_bench_late${late}_$i.json 2> /dev/null
done; done
tail -3 gpurun_out/g_pytest.log
for f in gpurun_out/g_bench_late*; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"; done
