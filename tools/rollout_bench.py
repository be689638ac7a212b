"""Closed-loop rollout throughput (BASELINE.json configs[4]): N robots x T ticks, cold and warm-started."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biped_mpc_py_b200 import BatchedMPC, MPC, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["cold", "warm"]
s = BatchedMPC(MPC(), synth.rollout_biped(), max_batch=n)
dev = s.device
b = synth.make_rollout_batch(n, shard_index=0)
tn = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
for mode in modes:
    st = [tn(b["x"], torch.float64), tn(b["foot"], torch.float64), tn(b["tick"], torch.int32), tn(b["gait"], torch.uint8),
          tn(b["q"], torch.float64), tn(b["qd"], torch.float64)]
    s.rollout(*[t.clone() for t in st], 3, warm_start=(mode == "warm"))  # warm-up (allocates the workspace)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = s.rollout(*st, ticks, warm_start=(mode == "warm"))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sn = BatchedMPC.rollout_stats(out["stats"])
    x = st[0].cpu().numpy()
    print(json.dumps(dict(mode=mode, robots=n, ticks=ticks, ms=ms, robot_ticks_per_s=n * ticks / (ms * 1e-3),
                          ms_per_tick=ms / ticks, **sn, z_mean=float(x[:, 5].mean()),
                          euler_absmax=float(np.abs(x[:, :3]).max()), finite=bool(np.isfinite(x).all()))), flush=True)
