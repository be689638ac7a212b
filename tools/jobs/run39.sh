cd $GRAFT_REPO_ROOT
timeout 600 python tools/lane_probe.py 262144 4:0:2:0 2>&1 | tail -1 | tee gpurun_out/probe27.log
SIZES="16384 32768 65536 131072 262144" bash tools/jobs/run34.sh 2>&1 | grep "^n="
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
