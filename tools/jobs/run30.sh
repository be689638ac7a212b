cd $GRAFT_REPO_ROOT
timeout 900 python tools/lane_probe.py 262144 4:2:2:0 2>&1 | tail -1 | tee gpurun_out/probe24.log
for lib in l2s l2ks l2s l2ks; do
EXP_LIB=$lib timeout 900 python tools/lane_probe.py 262144 4:2:2:0 2>&1 | tail -1 | tee -a gpurun_out/probe24.log
done
timeout 900 python tools/lane_probe.py 262144 4:2:2:0 2>&1 | tail -1 | tee -a gpurun_out/probe24.log
