cd $GRAFT_REPO_ROOT
timeout 900 python tools/lane_probe.py 262144 4:2:2:1,4:2:2:0,4:2:2:1,4:2:2:0 2>&1 | tee gpurun_out/probe20.log
