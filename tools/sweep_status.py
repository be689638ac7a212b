"""Robustness sweep: several shards of the synthetic distribution, count instances not certified optimal."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
shards = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [10, 11, 12, 13]
h = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mpc = MPC(h=h)
s = BatchedMPC(mpc, Biped(), max_batch=n, extend_gait=(h != 10))
tot = np.zeros(4, dtype=np.int64)
for sh in shards:
    b = synth.make_batch(n, shard_index=sh, mpc=mpc, extend=(h != 10))
    out = s.step_host(b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"], phase_k=b["phase_k"])
    c = np.bincount(out["status"], minlength=4)
    tot += c
    print("h", h, "shard", sh, "status", c.tolist(), "iters mean %.3f max %d" % (out["iters"].mean(), out["iters"].max()), flush=True)
print("TOTAL", tot.tolist())
