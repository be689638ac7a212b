cd $GRAFT_REPO_ROOT
python - <<'PY' 2>&1 | tail -10
import sys, numpy as np, torch, time
sys.path.insert(0, '.')
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = 65536
mpc = MPC(h=30)
s = BatchedMPC(mpc, Biped(), max_batch=n, extend_gait=True)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
for shard in range(1000, 1008):
    b = synth.make_batch(n, shard_index=shard, mpc=mpc, extend=True)
    d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8), tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
    s.step(*d); torch.cuda.synchronize()
    s.enable_timing(True); out = s.step(*d); ser = s.last_timing_ms(); s.enable_timing(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = s.step(*d); e1.record(); torch.cuda.synchronize()
    print(f"shard {shard}: tick {e0.elapsed_time(e1):.1f} ms; serial intervals (classify, lane walking, lane standing, warp walking, warp standing + re-solve + last resort) {[round(x,1) for x in ser]}; iters max {int(out['iters'].max())} status {np.bincount(out['status'].cpu().numpy(), minlength=4).tolist()}", flush=True)
PY
