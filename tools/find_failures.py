"""Run a synthetic batch on the GPU and dump the instances that did not reach status 0."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
shard = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mpc, biped = MPC(), Biped()
b = synth.make_batch(n, shard_index=shard, mpc=mpc, biped=biped)
s = BatchedMPC(mpc, biped, max_batch=n)
out = s.step_host(b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"])
bad = np.nonzero(out["status"] != 0)[0]
print("n", n, "status counts", np.bincount(out["status"], minlength=4).tolist(), "iters mean", out["iters"].mean(),
      "hist", np.bincount(out["iters"]).tolist())
np.savez("gpurun_out/failures.npz", idx=bad, status=out["status"][bad], iters=out["iters"][bad], resid=out["resid"][bad],
         gait=b["gait"][bad], controls=out["controls"][bad], n=n, shard=shard)
for i in bad[:20]:
    print(i, "gait", b["gait"][i], "status", out["status"][i], "iters", out["iters"][i], "resid", out["resid"][i])
