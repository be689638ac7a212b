"""Small fixed workload for compute-sanitizer: one tick through the default dispatch (warp-per-robot kernels: shared-memory
tile matrix, mbarriers, TMA bulk copies) and one with the lane-per-robot kernels and their later passes forced on, a short
warm-started closed-loop rollout, and the h = 30 kernels.   usage: python tools/san_driver.py [n] [what: warp|lane|rollout|h30|all]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
what = sys.argv[2] if len(sys.argv) > 2 else "all"


def tick(hz, force_lane):
    mpc, biped = MPC(h=hz), Biped()
    b = synth.make_batch(n, shard_index=3, mpc=mpc, biped=biped, extend=(hz != 10))
    s = BatchedMPC(mpc, biped, max_batch=n, extend_gait=(hz != 10))
    if force_lane:
        s.pin_kernel_family("lane")
    out = s.step_host(b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"], phase_k=b["phase_k"], want_states=True)
    print(f"h={hz} lane={force_lane}: status", np.bincount(out["status"], minlength=4).tolist(), "iters", float(out["iters"].mean()), flush=True)
    s.close()


if what in ("warp", "all"):
    tick(10, False)
if what in ("lane", "all"):
    tick(10, True)
if what in ("h30", "all"):
    tick(30, False)
if what in ("rollout", "all"):
    mpc, biped = MPC(), synth.rollout_biped()
    b = synth.make_batch(n, shard_index=4, mpc=mpc, biped=biped)
    s = BatchedMPC(mpc, biped, max_batch=n)
    dev = s.device
    tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    x = tn(np.tile(np.array([0, 0, 0, 0, 0, 0.55, 0, 0, 0, 0, 0, 0.0]), (n, 1)) + 0.02 * (b["x_fb"] - b["x_fb"].mean(axis=0)))
    r = s.rollout(x, tn(b["foot"]), torch.zeros(n, dtype=torch.int32, device=dev), tn(b["gait"], torch.uint8), tn(b["q"]), tn(b["qd"]),
                  ticks=6, warm_start=True)
    torch.cuda.synchronize()
    print("rollout stats", BatchedMPC.rollout_stats(r["stats"].cpu().numpy()) if "stats" in r else "", flush=True)
    s.close()
