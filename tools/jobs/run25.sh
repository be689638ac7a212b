cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_r2c.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r2c.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 300 gpurun_out/bench_r2c.json; echo
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2c_ref.json 2> gpurun_out/bench_r2c_ref.err; tail -c 200 gpurun_out/bench_r2c_ref.json; echo
bash tools/jobs/sanitizer.sh
