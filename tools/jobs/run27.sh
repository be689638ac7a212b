cd $GRAFT_REPO_ROOT
timeout 900 python tools/lane_probe.py 131072 4:2:2:0:16:1:0,4:2:2:0:16:0:0,4:2:2:0:16:1:18,4:2:2:0:16:2:0 30 2>&1 | tee gpurun_out/probe21.log
