"""Small fixed h=30 workload for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mpc = MPC(h=30)
b = synth.make_batch(n, shard_index=0, mpc=mpc, extend=True)
s = BatchedMPC(mpc, Biped(), max_batch=n, extend_gait=True)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8), tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
out = s.step(*d)
torch.cuda.synchronize()
print("status", np.bincount(out["status"].cpu().numpy(), minlength=4).tolist(), "iters", float(out["iters"].float().mean()))
