cd $GRAFT_REPO_ROOT
mkdir -p /tmp/ncu
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:lane_tick -c 6 -o /tmp/ncu/lane_v40_262144 -f python tools/prof_driver.py 262144 1 > gpurun_out/ncu_v40.log 2>&1; tail -2 gpurun_out/ncu_v40.log
ncu -i /tmp/ncu/lane_v40_262144.ncu-rep --page raw --csv > gpurun_out/lane_v40_raw.csv 2>/dev/null
ncu -i /tmp/ncu/lane_v40_262144.ncu-rep --page source --csv --print-source=sass,cuda 2>/dev/null | gzip -9 > gpurun_out/lane_v40_source.csv.gz
ls -la /tmp/ncu gpurun_out/lane_v40*
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_r2b.log 2>&1; tail -1 gpurun_out/ncu_launch_r2b.log | cut -c1-100
