cd $GRAFT_REPO_ROOT
timeout 900 python tools/gate_probe.py 10 1024,2048,4096,8192,16384,32768 > gpurun_out/gate10.log 2>&1; cat gpurun_out/gate10.log
timeout 900 python tools/gate_probe.py 30 512,1024,2048,4096,8192 > gpurun_out/gate30.log 2>&1; cat gpurun_out/gate30.log
