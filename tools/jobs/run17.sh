cd $GRAFT_REPO_ROOT
EXP_LIB=hoist2 timeout 900 python tools/lane_probe.py 262144 4:2:2:1,2:4:2:1,8:1:2:1,4:2:1:1,4:2:0:1,1:8:0:1,2:4:0:1,4:2:2:0 2>&1 | tee -a gpurun_out/probe12.log
