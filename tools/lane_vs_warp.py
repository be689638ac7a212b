"""Large-batch cross-check of the two kernel families: the same shards through the default path (lane-per-robot kernels
for both classes) and through the warp-per-robot kernels only (BMPC_LANE=0); reports certification and the largest
per-robot difference of the returned optimum.  usage: python tools/lane_vs_warp.py [n] [shard,shard,...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
shards = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [20, 21, 22, 23]
mpc, biped = MPC(), Biped()
os.environ.pop("BMPC_LANE", None)
lane = BatchedMPC(mpc, biped, max_batch=n)
os.environ["BMPC_LANE"] = "0"
warp = BatchedMPC(mpc, biped, max_batch=n)
worst_u = worst_tau = 0.0
tot = np.zeros(4, dtype=np.int64)
for sh in shards:
    b = synth.make_batch(n, shard_index=sh, mpc=mpc, biped=biped)
    args = (b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"])
    a = lane.step_host(*args, phase_k=b["phase_k"])
    l0 = lane.launch_count
    w = warp.step_host(*args, phase_k=b["phase_k"])
    tot += np.bincount(a["status"], minlength=4)
    assert (w["status"] == 0).all()
    scale = np.maximum(1.0, np.abs(w["controls"]).reshape(n, -1).max(axis=1))
    du = np.abs(a["controls"] - w["controls"]).reshape(n, -1).max(axis=1) / scale
    dt = np.abs(a["tau"] - w["tau"]).max(axis=1)
    worst_u, worst_tau = max(worst_u, float(du.max())), max(worst_tau, float(dt.max()))
    print(f"shard {sh}: lane-path status {np.bincount(a['status'], minlength=4).tolist()} iters {a['iters'].mean():.3f} | "
          f"max rel |du| {du.max():.3e} max |dtau| {dt.max():.3e} fric masks equal {(a['fric_active'] == w['fric_active']).mean():.6f}", flush=True)
print("TOTAL status", tot.tolist(), "worst rel |du|", worst_u, "worst |dtau|", worst_tau)
