cd $GRAFT_REPO_ROOT
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:lane_tick -c 2 -o gpurun_out/lane_v32_262144 -f python tools/prof_driver.py 262144 1 > gpurun_out/ncu_v32.log 2>&1; tail -2 gpurun_out/ncu_v32.log
