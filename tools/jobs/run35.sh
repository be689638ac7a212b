cd $GRAFT_REPO_ROOT
SIZES="16384 32768 65536 131072 262144" bash tools/jobs/run34.sh 2>&1 | grep "^n="
SIZES="65536 262144" bash tools/jobs/run34.sh 2>&1 | grep "^n="
timeout 600 python tools/lane_probe.py 65536 4:0:2:0,4:2:2:0 2>&1 | tail -2
