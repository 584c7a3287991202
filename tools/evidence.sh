#!/bin/bash
# Re-takes the judged measurement evidence on the CURRENT binary (one B200, run under gpurun):
#   tools/evidence.sh <tag>        e.g. r02a  ->  gpurun_out/<tag>_*  (copy what is to be judged into profiles/)
# Steps: build id, launch list + DRAM bytes of one UNet call (P = 64), per-shape event table, ncu --set full of the
# contraction / GroupNorm kernels at four places of the call, ncu DRAM bytes of the DWT at B = 256 (1.2 GB >> L2).
# ncu numbers are cold-cache and serialised: shares and byte counts are the evidence, bench.py's CUDA events the timings.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
sha256sum wavedm_b200/libwavedm_b200.so | cut -c1-16 > $O/${TAG}_lib_sha16.txt
python -c "from wavedm_b200 import _lib; print(_lib.source_id())" > $O/${TAG}_src_sha16.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/${TAG}_smi.txt 2>&1

# 1. launch list + DRAM bytes of one UNet call
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file $O/${TAG}_dram_unet_p64.csv python tools/profile_unet.py --patches 64 --iters 1 > $O/${TAG}_ncu1.log 2>&1
python - "$O/${TAG}_dram_unet_p64.csv" > $O/${TAG}_launches_unet_p64.csv <<'EOF'
import csv, sys
rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
r = list(csv.DictReader(rows))
w = csv.DictWriter(sys.stdout, fieldnames=r[0].keys())
w.writeheader()
for x in r:
    if x["Metric Name"] == "gpu__time_duration.sum":
        w.writerow(x)
EOF
python tools/summarize_launches.py $O/${TAG}_launches_unet_p64.csv > $O/${TAG}_launch_summary.txt 2>&1
python tools/traffic_from_ncu.py $O/${TAG}_dram_unet_p64.csv > $O/${TAG}_traffic.json 2> $O/${TAG}_traffic.err
python - "$O/${TAG}_traffic.json" "$O/${TAG}_lib_sha16.txt" <<'EOF'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    d["lib_sha16"] = open(sys.argv[2]).read().strip()
    d["src_sha16"] = open(sys.argv[2].replace("_lib_sha16", "_src_sha16")).read().strip()
    json.dump(d, open(sys.argv[1], "w"), indent=1)
except Exception as e:
    print("traffic json:", e)
EOF

# 2. per-shape event table (CUDA events on the launch stream)
timeout 300 python tools/profile_unet.py --patches 64 --iters 5 --time --spans > $O/${TAG}_spans_events.txt 2>&1

# 3. ncu --set full at four places of the call: level-0 down path, 32x32 / 16x16 (+attention), 8x8 mid, level-0 up path + conv_out
i=0
for skip in 8 40 100 160; do
    i=$((i + 1))
    timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc|gn_apply|gn_finalize" -s $skip -c 10 -f \
        -o $O/${TAG}_full_$i python tools/profile_unet.py --patches 64 --iters 1 > $O/${TAG}_ncu_full_$i.log 2>&1
    python tools/ncu_summary.py $O/${TAG}_full_$i.ncu-rep > $O/${TAG}_ncu_full_$i.csv 2>> $O/${TAG}_ncu_full_$i.log
    rm -f $O/${TAG}_full_$i.ncu-rep   # gpurun_out/ is capped at 64 MiB: only the extracted summaries travel back
done

# 4. DWT / IWT at B = 256 (3+ x 403 MB rotating buffers): DRAM bytes per launch next to the algorithmic 402 653 184 B.
#    bench_dwt.py launches every kernel 45 times per shape, B = 64 first: skipping 50 launches of one kernel lands in B = 256
: > $O/${TAG}_dwt_dram.csv
for k in dwt4x4_direct iwt4x4_direct; do
    timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:$k \
        -s 50 -c 8 --csv --log-file $O/${TAG}_dwt_dram_$k.csv python tools/bench_dwt.py > $O/${TAG}_dwt_ncu.log 2>&1
    grep -v "^==" $O/${TAG}_dwt_dram_$k.csv >> $O/${TAG}_dwt_dram.csv
    rm -f $O/${TAG}_dwt_dram_$k.csv
done
timeout 300 python tools/bench_dwt.py > $O/${TAG}_bench_dwt.txt 2>&1

# 5. HFRM engine: per-kernel launch list of one call (B = 64, 256x256, bf16) + timing
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"hfrm_|gemm_tc" -c 700 --csv \
    --log-file $O/${TAG}_hfrm_launches.csv python tools/bench_hfrm.py --precisions bf16 --iters 1 > $O/${TAG}_hfrm_ncu.log 2>&1
python tools/hfrm_launch_summary.py $O/${TAG}_hfrm_launches.csv > $O/${TAG}_hfrm_launch_summary.txt 2>&1
timeout 300 python tools/bench_hfrm.py > $O/${TAG}_bench_hfrm.txt 2>&1
# 6. per-step GPU timeline of the sampler + single-image latency
timeout 300 python tools/sampler_timeline.py 2>&1 | tail -6 > $O/${TAG}_sampler_timeline.txt
timeout 300 python tools/latency_small.py > $O/${TAG}_latency_small.txt 2>&1
# 7. single-image call (P = 1): per-shape event table with split-K, and the training-step update kernel
timeout 300 python tools/profile_unet.py --patches 1 --iters 5 --time --spans > $O/${TAG}_p1_spans_splitk.txt 2>&1
WDM_TC_SPLITK=0 timeout 300 python tools/latency_small.py > $O/${TAG}_latency_small_nosplit.txt 2>&1
timeout 600 python tools/bench_train_step.py --json $O/${TAG}_train_step.json > $O/${TAG}_train_step.txt 2>&1
du -sh $O
ls -la $O | grep ${TAG}_ | head -40
