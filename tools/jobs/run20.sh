cd $GRAFT_REPO_ROOT
# cfg: warps:ctas:sync:prefetch:polish_rounds:inline_rounds:ipm_inline
EXP_LIB=park4 timeout 900 python tools/lane_probe.py 262144 4:2:2:1:4:0:0,4:2:2:1:4:1:0,4:2:2:1:4:1:10,4:2:2:1:4:1:9,4:2:2:1:4:1:11,4:2:2:1:4:1:8 2>&1 | tee -a gpurun_out/probe17.log
EXP_LIB=park4 timeout 900 python tools/lane_probe.py 32768 4:2:2:1:4:0:0,4:2:2:1:4:1:10 2>&1 | tee -a gpurun_out/probe17.log
