cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_r2g.log
bash tools/jobs/run44.sh 2>&1 | grep "shard" | cut -c1-60
