cd $GRAFT_REPO_ROOT
timeout 900 python tools/lane_probe.py 262144 4:2:2:0 2>&1 | tail -1 | tee gpurun_out/probe25.log
for i in 1 2; do
EXP_LIB=f32sl timeout 900 python tools/lane_probe.py 262144 4:2:2:0 2>&1 | tail -1 | tee -a gpurun_out/probe25.log
done
timeout 900 python tools/lane_probe.py 262144 4:2:2:0 2>&1 | tail -1 | tee -a gpurun_out/probe25.log
EXP_LIB=f32sl timeout 900 python tools/lane_probe.py 32768 4:2:2:0 2>&1 | tail -1 | tee -a gpurun_out/probe25.log
