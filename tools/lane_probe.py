"""Throughput probe of the lane-per-robot kernels (device-resident inputs): per-class kernel times for a set of
residency settings, and agreement with the warp-per-robot kernels on a sample.
usage: python tools/lane_probe.py [n] [warps:ctas[:sync[:prefetch]],...] [h]     (warps = 0: warp-per-robot kernels only)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if os.environ.get("EXP_LIB"):  # an experiment library built by tools/exp_build.sh
    import biped_mpc_py_b200._lib as _l
    _l.LIB_PATH = os.path.join(os.path.dirname(_l.LIB_PATH), "_exp", f"lib_{os.environ['EXP_LIB']}.so")
    print("library:", _l.LIB_PATH, flush=True)
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
cfgs = [tuple(int(x) for x in c.split(":")) for c in (sys.argv[2] if len(sys.argv) > 2 else "4:0").split(",")]
hz = int(sys.argv[3]) if len(sys.argv) > 3 else 10
reps = int(os.environ.get("REPS", "3"))
mpc, biped = MPC(h=hz), Biped()
b = synth.make_batch(n, shard_index=0, mpc=mpc, biped=biped, extend=(hz != 10))
s = BatchedMPC(mpc, biped, max_batch=n, extend_gait=(hz != 10))
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8),
     tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
walk = int((b["contact"].astype(int).sum(axis=(1, 2)) <= hz).sum())
print(f"n={n} h={hz} walking {walk} standing {n - walk}", flush=True)
s.enable_timing(True)
ref = None
for cfg in cfgs:
    warps, ctas = cfg[0], cfg[1]
    sync = cfg[2] if len(cfg) > 2 else 2
    pref = cfg[3] if len(cfg) > 3 else 1
    if warps == 0:
        s.set_option("lane_mode", 0)
    else:
        s.set_option("lane_mode", 2)
        s.set_option("lane_min", 1)
        s.set_option("lane_warps", warps)
        s.set_option("lane_ctas_per_sm", ctas)
        s.set_option("lane_sync", sync)
        s.set_option("lane_prefetch", pref)
        if len(cfg) > 4:
            s.set_option("polish_rounds", cfg[4])
        if len(cfg) > 5:
            s.set_option("lane_inline_rounds", cfg[5])
        if len(cfg) > 6:
            s.set_option("lane_ipm_inline", cfg[6])
        if len(cfg) > 7:
            s.set_option("lane_ctas_standing", cfg[7])
    out = s.step(*d)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        out = s.step(*d)
        ts.append(s.last_timing_ms())
    ts = np.array(ts).min(axis=0)
    s.enable_timing(False)  # the real thing: the two class chains run concurrently
    tot = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = s.step(*d)
        e1.record()
        torch.cuda.synchronize()
        tot = min(tot, e0.elapsed_time(e1))
    s.enable_timing(True)
    st = np.bincount(out["status"].cpu().numpy(), minlength=4).tolist()
    it = out["iters"].cpu().numpy()
    line = (f"warps/cta {warps} ctas/sm {ctas} sync {sync} prefetch {pref}: lane {ts[1]:.2f} + {ts[2]:.2f} ms warp-per-robot {ts[3]:.2f} + {ts[4]:.2f} ms -> "
            f"serial sum {ts.sum():.2f} ms, concurrent {tot:.2f} ms = {n / (tot * 1e-3) / 1e6:.3f} M solves/s  status {st} iters {it.mean():.3f}")
    u = out["controls"].cpu().numpy()
    if ref is None:
        ref = u
    else:
        sc = np.maximum(1.0, np.abs(ref).reshape(n, -1).max(axis=1))
        line += f"  max rel diff vs first cfg {(np.abs(u - ref).reshape(n, -1).max(axis=1) / sc).max():.2e}"
    print(line, flush=True)
