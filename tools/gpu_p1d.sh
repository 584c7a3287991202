mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,launch__grid_size --clock-control none -k regex:"gemm_tc|splitk" -c 200 --csv --log-file gpurun_out/p1d_launches.csv python tools/profile_unet.py --patches 1 --iters 1 > gpurun_out/p1d.log 2>&1
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/p1d_launches.csv') if not l.startswith('==')]
r=list(csv.DictReader(rows))
by={}
for x in r:
    by.setdefault(x['ID'],{'k':x['Kernel Name'][:48]})[x['Metric Name']]=x['Metric Value']
for i,(k,v) in enumerate(by.items()):
    if i>=60 and i<150: print(k, v)
PY
