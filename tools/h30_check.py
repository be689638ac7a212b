"""h = 30: the instance no kernel certified in round 1 (shard 1000, index 6464, standing; nearly degenerate, its polish cycled
between releasing and re-adding rows) alone and inside its 65,536-instance batch, through the default dispatch and through
the warp-per-robot kernels only; plus more shards, to count what is left uncertified."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = 65536
mpc = MPC(h=30)
for shard in (1000, 1001, 1002, 1003):
    b = synth.make_batch(n, shard_index=shard, mpc=mpc, extend=True)
    for mode in (("default", None), ("warp-per-robot only", 0)) if shard == 1000 else (("default", None),):
        s = BatchedMPC(mpc, Biped(), max_batch=n, extend_gait=True)
        if mode[1] is not None:
            s.set_option("lane_mode", mode[1])
        out = s.step_host(b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"], phase_k=b["phase_k"])
        bad = np.nonzero(out["status"] != 0)[0]
        print(f"shard {shard} {mode[0]}: status {np.bincount(out['status'], minlength=4).tolist()} bad idx {bad.tolist()} "
              f"gait {b['gait'][bad].tolist()} iters {out['iters'][bad].tolist()} max iters {int(out['iters'].max())}", flush=True)
        if shard == 1000 and mode[1] is None:
            one = s.step_host(*[b[k][6464:6465] for k in ("x_fb", "t", "foot", "contact", "q", "qd", "pf_w")], phase_k=b["phase_k"][6464:6465])
            print("  instance 6464 alone: status", one["status"].tolist(), "iters", one["iters"].tolist(),
                  "max |du| vs in-batch", float(np.abs(one["controls"][0] - out["controls"][6464]).max()))
        s.close()
