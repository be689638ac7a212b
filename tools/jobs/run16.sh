cd $GRAFT_REPO_ROOT
for lib in f4 hoist; do
EXP_LIB=$lib timeout 600 python tools/lane_probe.py 262144 4:2:2:1,4:2:10:1 2>&1 | tee -a gpurun_out/probe11.log
done
EXP_LIB=hoist timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or config1 or lane_per" 2>&1 | tail -3
