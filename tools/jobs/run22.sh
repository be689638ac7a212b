cd $GRAFT_REPO_ROOT
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 600 gpurun_out/bench_r2b.json; tail -3 gpurun_out/bench_r2b.err
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2b_ref.json 2> gpurun_out/bench_r2b_ref.err; tail -c 300 gpurun_out/bench_r2b_ref.json
