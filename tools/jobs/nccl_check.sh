cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-extra > gpurun_out/nccl_n2.json 2> gpurun_out/nccl_n2.err
grep -c "NCCL INFO" gpurun_out/nccl_n2.err; grep "NCCL INFO" gpurun_out/nccl_n2.err | grep -i "nranks\|NVLS\|Init COMPLETE" | head -6 | cut -c1-250
tail -c 200 gpurun_out/nccl_n2.json
