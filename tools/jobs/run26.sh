cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_rollout.py -m gpu -x -q -k "growing" 2>&1 | tail -30
