cd $GRAFT_REPO_ROOT
timeout 900 python tools/e2e_probe.py 262144 1,2,4 > gpurun_out/e2e1.log 2>&1; cat gpurun_out/e2e1.log
