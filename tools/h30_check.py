"""h = 30, 65,536 instances of shard 1000 (the batch of bench.py's h30 leg on rank 0) through the default path and through the
warp-per-robot kernels only: which instances are not certified.  Round 1: instance 6464 (standing) returns status 1 on BOTH paths."""
import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = 65536
mpc = MPC(h=30)
b = synth.make_batch(n, shard_index=1000, mpc=mpc, extend=True)
res = {}
for mode in ("2", "0"):
    os.environ["BMPC_LANE"] = mode
    s = BatchedMPC(mpc, Biped(), max_batch=n, extend_gait=True)
    out = s.step_host(b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"], phase_k=b["phase_k"])
    res[mode] = out
    bad = np.nonzero(out["status"] != 0)[0]
    print("BMPC_LANE", mode, "status", np.bincount(out["status"], minlength=4).tolist(), "bad idx", bad.tolist(), "gait", b["gait"][bad].tolist(), "iters", out["iters"][bad].tolist(), flush=True)
    s.close()
