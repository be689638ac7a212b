cd $GRAFT_REPO_ROOT
python tools/h30_check.py 2>&1 | tail -8
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
