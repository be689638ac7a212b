cd $GRAFT_REPO_ROOT
EXP_LIB=defer timeout 900 python tools/lane_probe.py 262144 4:2:2:1:4:0,4:2:2:1:4:1,4:2:2:1:4:2 2>&1 | tee -a gpurun_out/probe14.log
EXP_LIB=defer timeout 900 python tools/lane_probe.py 32768 4:2:2:1:4:0,4:2:2:1:4:1 2>&1 | tee -a gpurun_out/probe14.log
EXP_LIB=defer timeout 900 python tools/lane_probe.py 65536 4:2:2:1:16:0,4:2:2:1:16:1 30 2>&1 | tee -a gpurun_out/probe14.log
