cd $GRAFT_REPO_ROOT
timeout 600 python tools/lane_probe.py 262144 4:2:2:0 2>&1 | tail -1 | tee gpurun_out/probe26.log
for i in 1 2; do
EXP_LIB=tma timeout 300 python tools/lane_probe.py 262144 4:2:2:0 2>&1 | tail -1 | tee -a gpurun_out/probe26.log
done
EXP_LIB=tma timeout 300 python tools/lane_probe.py 32768 4:2:2:0 2>&1 | tail -1 | tee -a gpurun_out/probe26.log
