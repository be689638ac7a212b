"""Quick throughput check of the tick kernels (device-resident inputs): prints ms per class kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mpc, biped = MPC(), Biped()
b = synth.make_batch(n, shard_index=0, mpc=mpc, biped=biped)
mu_tol = float(os.environ.get('BMPC_MU_TOL', '0'))
rd_tol = float(os.environ.get('BMPC_RD_TOL', '0'))
s = BatchedMPC(mpc, biped, max_batch=n, mu_tol=mu_tol, rd_tol=rd_tol)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8),
     tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
for _ in range(2):
    out = s.step(*d)
s.enable_timing(True)
ts = []
for _ in range(4):
    out = s.step(*d)
    ts.append(s.last_timing_ms())
ts = np.array(ts).mean(axis=0)
st = np.bincount(out["status"].cpu().numpy(), minlength=4).tolist()
print(f"mu_tol={mu_tol} rd_tol={rd_tol} n={n} NWW={os.environ.get('BMPC_NW_WALK','8')} NTS={os.environ.get('BMPC_NT_STAND','128')} "
      f"lane {ts[1]:.2f} + {ts[2]:.2f} warp-per-robot {ts[3]:.2f} + {ts[4]:.2f} ms -> {n/(ts.sum()*1e-3)/1e6:.3f} M solves/s  status {st} "
      f"iters {float(out['iters'].float().mean()):.3f}")
