cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or config1 or lane_per" 2>&1 | tail -3
timeout 600 python tools/lane_probe.py 262144 4:2:2:1,4:2:6:1 > gpurun_out/probe10.log 2>&1; cat gpurun_out/probe10.log
timeout 600 python tools/lane_probe.py 32768 4:2:2:1 2>&1 | tail -1
