"""Horizon-30 throughput (BASELINE.json configs[3]): device-resident inputs, per-class kernel times."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mpc = MPC(h=30)
b = synth.make_batch(n, shard_index=0, mpc=mpc, extend=True)
s = BatchedMPC(mpc, Biped(), max_batch=n, extend_gait=True)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8),
     tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
for _ in range(2):
    out = s.step(*d)
s.enable_timing(True)
ts = []
for _ in range(3):
    out = s.step(*d)
    ts.append(s.last_timing_ms())
ts = np.array(ts).mean(axis=0)
st = np.bincount(out["status"].cpu().numpy(), minlength=4).tolist()
it = out["iters"].cpu().numpy()
print(json.dumps(dict(h=30, n=n, walking=int(b["gait"].sum()), standing=int(n - b["gait"].sum()), lane_walking_ms=float(ts[1]), lane_standing_ms=float(ts[2]),
                      warp_walking_ms=float(ts[3]), warp_standing_ms=float(ts[4]), solves_per_s=n / (ts.sum() * 1e-3), status=st, mean_iters=float(it.mean()),
                      max_iters=int(it.max()))))
