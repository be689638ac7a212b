cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_r2e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r2e.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; tail -c 200 gpurun_out/bench_r2e.json; echo
