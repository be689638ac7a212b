cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_r2f.log
bash tools/jobs/run42.sh 2>&1 | grep "shard" | cut -c1-160
timeout 600 python tools/lane_probe.py 262144 4:0:2:0 2>&1 | tail -1
