"""Crossover between the lane-per-robot and the warp-per-robot kernels for small classes: pure walking / pure standing
batches of increasing size, each through both paths (device-resident inputs, best of a few repetitions).
usage: python tools/gate_probe.py [h] [sizes,...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

hz = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1024,2048,4096,8192,16384,32768").split(",")]
mpc, biped = MPC(h=hz), Biped()
for wp, name in ((1.0, "walking"), (0.0, "standing")):
    for n in sizes:
        b = synth.make_batch(n, shard_index=3, mpc=mpc, biped=biped, walking_prob=wp, extend=(hz != 10))
        s = BatchedMPC(mpc, biped, max_batch=n, extend_gait=(hz != 10))
        dev = s.device
        tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
        d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8),
             tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
        res = {}
        for mode in ("warp", "lane"):
            if mode == "warp":
                s.set_option("lane_mode", 0)
            else:
                s.set_option("lane_mode", 2)
                s.set_option("lane_min", 1)
            s.step(*d)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = s.step(*d)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[mode] = best
        print(f"h={hz} {name:8s} n={n:6d}: warp-per-robot {res['warp']:7.2f} ms  lane {res['lane']:7.2f} ms  "
              f"status {np.bincount(out['status'].cpu().numpy(), minlength=2).tolist()}", flush=True)
        s.close()
