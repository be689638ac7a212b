set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or lane_per_robot or library" 2>&1 | tail -3
timeout 900 python tools/lane_probe.py 262144 4:2:2:1,8:1:1:1,8:1:2:1 > gpurun_out/probe6.log 2>&1; cat gpurun_out/probe6.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_tick -c 1 -o gpurun_out/lane_v25 python tools/lane_probe.py 65536 4:2:2:1 > gpurun_out/ncu_v25.log 2>&1; tail -3 gpurun_out/ncu_v25.log
