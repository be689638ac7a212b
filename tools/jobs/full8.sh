cd $GRAFT_REPO_ROOT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_n8_full.json 2> gpurun_out/bench_r2_n8_full.err
tail -1 gpurun_out/bench_r2_n8_full.json | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('n8 value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))
for k,v in d['other_configs'].items(): print(k, {a:(round(b) if isinstance(b,float) and b>1000 else b) for a,b in v.items() if a!='workload'})"
grep -c "NCCL INFO" gpurun_out/bench_r2_n8_full.err; grep "nranks 8" gpurun_out/bench_r2_n8_full.err | head -2 | cut -c1-160
