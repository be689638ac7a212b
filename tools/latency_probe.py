"""N=1 latency breakdown: end-to-end pinned tick vs device time of the three kernels."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

mpc, biped = MPC(), Biped()
for gait_sel in (1, 0):
    b = synth.make_batch(64, shard_index=3)
    i = int(np.nonzero(b["gait"] == gait_sel)[0][0])
    one = BatchedMPC(mpc, biped, max_batch=1)
    t1 = one.pinned_tick(1)
    for k in ("x_fb", "foot", "q", "qd", "pf_w", "t", "phase_k", "contact"):
        t1.inputs[k][...] = b[k][i:i + 1]
    lat = []
    for _ in range(300):
        a = time.perf_counter(); t1.run(); lat.append(time.perf_counter() - a)
    lat = np.array(lat[50:]) * 1e3
    one.enable_timing(True)
    ks = []
    for _ in range(50):
        t1.run(); ks.append(one.last_timing_ms())
    ks = np.median(np.array(ks), axis=0)
    print(f"gait={gait_sel} p50 {np.percentile(lat,50):.3f} ms p99 {np.percentile(lat,99):.3f} ms | device: classify {ks[0]*1e3:.1f} us, "
          f"walking kernel {ks[1]*1e3:.1f} us, standing kernel {ks[2]*1e3:.1f} us | iters {int(t1.outputs['iters'][0])}")
    one.close()
