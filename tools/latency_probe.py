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
          f"walking kernel {ks[3]*1e3:.1f} us, standing kernel {ks[4]*1e3:.1f} us | iters {int(t1.outputs['iters'][0])}")
    one.close()


# ---- a caller-owned control loop for ONE robot (states replayed from a device rollout), cold vs warm-started ticks ----
for gait_sel in (1, 0):
    rb = synth.make_rollout_batch(64, shard_index=4)
    i = int(np.nonzero(rb["gait"] == gait_sel)[0][0])
    ticks = 200
    s = BatchedMPC(MPC(), synth.rollout_biped(), max_batch=1)
    dev = s.device
    tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    st = [tn(rb[k][i:i + 1], dt) for k, dt in (("x", torch.float64), ("foot", torch.float64), ("tick", torch.int32), ("gait", torch.uint8),
                                                 ("q", torch.float64), ("qd", torch.float64))]
    log = s.rollout(*st, ticks, warm_start=False, n_log=1)
    torch.cuda.synchronize()
    xl, fl = log["x_log"].cpu().numpy()[:, 0], log["foot_log"].cpu().numpy()[:, 0]
    t1 = s.pinned_tick(1)
    for warm in (False, True):
        s.warm_start(warm)
        lat, its = [], []
        for k in range(ticks):
            T = int(rb["tick"][i]) + k
            rows = (T + np.arange(10)) % 10
            contact = np.stack([rows < 5, rows >= 5], axis=1) if gait_sel == 1 else np.ones((10, 2), bool)
            t1.inputs["x_fb"][0], t1.inputs["foot"][0], t1.inputs["pf_w"][0] = xl[k], fl[k], fl[k]
            t1.inputs["q"][0], t1.inputs["qd"][0], t1.inputs["t"][0] = rb["q"][i], rb["qd"][i], T * 0.04
            t1.inputs["phase_k"][0], t1.inputs["contact"][0] = T % 10, contact.astype(np.uint8)
            a = time.perf_counter(); out = t1.run(); lat.append(time.perf_counter() - a)
            its.append(int(out["iters"][0]))
            assert int(out["status"][0]) == 0
        lat = np.array(lat[20:]) * 1e3
        print(f"control loop, one robot, gait={gait_sel}, warm_start={warm}: p50 {np.percentile(lat,50):.3f} ms p99 {np.percentile(lat,99):.3f} ms, "
              f"mean interior-point iterations {np.mean(its[20:]):.2f}")
    s.close()
