cd $GRAFT_REPO_ROOT
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo rc=$?; tail -3 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lane_tick -c 2 -o gpurun_out/lane_v26_262144 python tools/prof_driver.py 262144 1 > gpurun_out/ncu_v26.log 2>&1; tail -2 gpurun_out/ncu_v26.log
