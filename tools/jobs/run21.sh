cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_r2b.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_r2b.log
timeout 600 python tools/lane_probe.py 262144 4:2:2:1 2>&1 | tee gpurun_out/probe18.log
timeout 600 python tools/lane_probe.py 32768 4:2:2:1 2>&1 | tail -1 | tee -a gpurun_out/probe18.log
timeout 600 python tools/lane_probe.py 65536 4:2:2:1:16:1:0,4:2:2:1:16:0:0,4:2:2:1:16:1:18,4:2:2:1:16:1:20 30 2>&1 | tee -a gpurun_out/probe18.log
