"""End-to-end host path: pinned host -> device -> kernels -> pinned host, unchunked vs chunked (copies of neighbouring chunks
under the kernels), against the device-resident time of the same batch.
usage: python tools/e2e_probe.py [n] [chunk counts, e.g. 1,2,4,8]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
counts = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
reps = 5
mpc, biped = MPC(), Biped()
b = synth.make_batch(n, shard_index=0, mpc=mpc, biped=biped)
s = BatchedMPC(mpc, biped, max_batch=n)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8),
     tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
for _ in range(2):
    s.step(*d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = s.step(*d)
e1.record()
torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / reps
ref_u = out["controls"].cpu().numpy()
print(f"n={n}: device-resident {dev_ms:.2f} ms per step = {n / dev_ms / 1e3:.3f} M solves/s", flush=True)
keys = ("x_fb", "foot", "q", "qd", "pf_w", "t", "phase_k", "contact")
for c in counts:
    if c == 1:
        tick = s.pinned_tick(n)
        for k in keys:
            tick.inputs[k][...] = b[k]
        run = tick.run
        first = lambda r: r
    else:
        tick = s.chunked_tick(n, c)
        tick.set_inputs(**{k: b[k] for k in keys})
        run = tick.run
        first = lambda r: r[0]
    for _ in range(2):
        run()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = run()
        _ = float(first(r)["tau"][0, 0])
    ms = (time.perf_counter() - t0) / reps * 1e3
    u = tick.outputs["controls"] if c == 1 else tick.gather()["controls"]
    same = bool(np.array_equal(u, ref_u)) if c == 1 else float(np.abs(u - ref_u).max())
    print(f"  chunks {c}: {ms:.2f} ms per step = {n / ms / 1e3:.3f} M solves/s ({dev_ms / ms:.3f} of device-resident), "
          f"H2D {tick.h2d_bytes / 1e6:.1f} MB D2H {tick.d2h_bytes / 1e6:.1f} MB, vs device-resident result: {same}", flush=True)

# consecutive steps pipelined over two slots: step i+1 is launched before step i's result is read on the host
for c in counts:
    ticks = [s.chunked_tick(n, c, slot=k) for k in range(2)]
    for t in ticks:
        t.set_inputs(**{k: b[k] for k in keys})
        t.run()
    steps = 2 * reps
    t0 = time.perf_counter()
    ticks[0].launch()
    for i in range(1, steps):
        ticks[i % 2].launch()
        r = ticks[(i - 1) % 2].wait()
        _ = float(r[0]["tau"][0, 0])
    r = ticks[(steps - 1) % 2].wait()
    _ = float(r[0]["tau"][0, 0])
    ms = (time.perf_counter() - t0) / steps * 1e3
    err = float(np.abs(ticks[1].gather()["controls"] - ref_u).max())
    print(f"  2 slots x chunks {c}: {ms:.2f} ms per step = {n / ms / 1e3:.3f} M solves/s ({dev_ms / ms:.3f} of device-resident), "
          f"max |diff| vs device-resident result {err}", flush=True)
