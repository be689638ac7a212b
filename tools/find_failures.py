"""Run a synthetic batch on the GPU and dump the instances that did not reach status 0."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
shard = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mpc, biped = MPC(h=h), Biped()
b = synth.make_batch(n, shard_index=shard, mpc=mpc, biped=biped, extend=(h != 10))
s = BatchedMPC(mpc, biped, max_batch=n, extend_gait=(h != 10))
out = s.step_host(b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"], phase_k=b["phase_k"])
bad = np.nonzero(out["status"] != 0)[0]
print("n", n, "status counts", np.bincount(out["status"], minlength=4).tolist(), "iters mean", out["iters"].mean(),
      "hist", np.bincount(out["iters"]).tolist())
np.savez(f"gpurun_out/failures_h{h}.npz", idx=bad, status=out["status"][bad], iters=out["iters"][bad], resid=out["resid"][bad],
         gait=b["gait"][bad], controls=out["controls"][bad], n=n, shard=shard)
for i in bad[:20]:
    print(i, "gait", b["gait"][i], "status", out["status"][i], "iters", out["iters"][i], "resid", out["resid"][i])
