cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1500 python bench.py --steps 5 --warmup 2 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo rc=$?; tail -3 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
