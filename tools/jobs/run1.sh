set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or lane_per_robot or foot_positions or library" 2>&1 | tail -15
timeout 600 python tools/lane_probe.py 262144 0:0,4:2,4:1,2:2,2:1,3:2,6:1 > gpurun_out/probe1.log 2>&1; cat gpurun_out/probe1.log
timeout 300 python tools/lane_probe.py 32768 0:0,4:2,4:1 > gpurun_out/probe1_32k.log 2>&1; cat gpurun_out/probe1_32k.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_tick -c 2 -o gpurun_out/lane_v20 python tools/lane_probe.py 65536 4:2 > gpurun_out/ncu_v20.log 2>&1; tail -3 gpurun_out/ncu_v20.log
