#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) >> $O
timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_g.json 2>> $O
python - >> $O <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_g.json') if l.startswith('{')][-1])
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['share_of_step'])
PY
cat $O
