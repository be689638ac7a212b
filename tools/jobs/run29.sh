cd $GRAFT_REPO_ROOT
timeout 900 python tools/lane_probe.py 262144 4:2:2:0:4:1:9:2,4:2:2:0:4:1:9:1,4:2:2:0:4:1:9:2,4:2:2:0:4:1:9:1 2>&1 | tee gpurun_out/probe22.log
timeout 900 python tools/lane_probe.py 131072 4:2:2:0:4:1:9:2,4:2:2:0:4:1:9:1 2>&1 | tee -a gpurun_out/probe22.log
timeout 900 python tools/lane_probe.py 65536 4:2:2:0:4:1:9:2,4:2:2:0:4:1:9:1 2>&1 | tee -a gpurun_out/probe22.log
