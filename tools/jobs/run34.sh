cd $GRAFT_REPO_ROOT
for n in ${SIZES:-1024 4096 8192 16384 32768 65536 131072 262144}; do
python - $n <<'PY' 2>&1 | tail -1 | tee -a gpurun_out/size_sweep_r2.log
import sys, numpy as np, torch
sys.path.insert(0, '.')
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = int(sys.argv[1])
mpc, biped = MPC(), Biped()
b = synth.make_batch(n, shard_index=0, mpc=mpc, biped=biped)
s = BatchedMPC(mpc, biped, max_batch=n)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8), tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
for _ in range(3): out = s.step(*d)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = s.step(*d); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
st = np.bincount(out["status"].cpu().numpy(), minlength=4).tolist()
print(f"n={n}: {best:.3f} ms per tick = {n / best / 1e3:.3f} M solves/s, status {st}")
PY
done
