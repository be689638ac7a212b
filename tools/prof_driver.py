"""Small fixed workload for ncu: a few ticks over a modest synthetic batch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mpc, biped = MPC(), Biped()
b = synth.make_batch(n, shard_index=0, mpc=mpc, biped=biped)
s = BatchedMPC(mpc, biped, max_batch=n)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8),
     tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
for _ in range(reps):
    out = s.step(*d)
torch.cuda.synchronize()
print("status", np.bincount(out["status"].cpu().numpy(), minlength=4).tolist(), "iters", float(out["iters"].float().mean()))
