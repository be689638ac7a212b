# usage (on the GPU box, N GPUs): bash tools/jobs/scale.sh N   -> gpurun_out/scale_r2_{weak,strong}_nN.json
cd $GRAFT_REPO_ROOT
N=$1
for mode in strong weak; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --scaling $mode --no-cpu-baseline --no-latency --no-extra > gpurun_out/scale_r2_${mode}_n1.json 2> gpurun_out/scale_r2_${mode}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --scaling $mode --no-cpu-baseline --no-latency --no-extra > gpurun_out/scale_r2_${mode}_n$N.json 2> gpurun_out/scale_r2_${mode}_n$N.err
  fi
  tail -1 gpurun_out/scale_r2_${mode}_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))"
done
