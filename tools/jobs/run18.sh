cd $GRAFT_REPO_ROOT
EXP_LIB=pr timeout 900 python tools/lane_probe.py 262144 4:2:2:1:4,4:2:2:1:1,4:2:2:1:2 2>&1 | tee -a gpurun_out/probe13.log
