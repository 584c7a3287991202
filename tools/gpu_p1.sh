mkdir -p gpurun_out
timeout 300 python tools/profile_unet.py --patches 1 --iters 5 --time --spans > gpurun_out/p1_spans.txt 2>&1
timeout 300 python tools/latency_small.py 2>&1 | grep "B=1 1024\|B=1 256x256 (1 patches/step) 50" > gpurun_out/p1_lat_base.txt
WDM_GN_FUSED_FINALIZE=1 timeout 300 python tools/latency_small.py 2>&1 | grep "B=1 256x256 (1 patches/step) 50" > gpurun_out/p1_lat_gnfused.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p1_launches.csv python tools/profile_unet.py --patches 1 --iters 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/p1_launches.csv > gpurun_out/p1_launch_summary.txt 2>&1
cat gpurun_out/p1_lat_base.txt gpurun_out/p1_lat_gnfused.txt; head -30 gpurun_out/p1_launch_summary.txt
