cd $GRAFT_REPO_ROOT
python - <<'PY' 2>&1 | tail -8
import sys, numpy as np, torch
sys.path.insert(0, '.')
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = 65536
mpc, biped = MPC(), Biped()
b = synth.make_batch(n, shard_index=0, mpc=mpc, biped=biped)
s = BatchedMPC(mpc, biped, max_batch=n)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8), tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
def times(k=8):
    ts=[]
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = s.step(*d); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1),2))
    return ts
s.enable_timing(True)
s.step(*d); torch.cuda.synchronize()
print('first steps with timing on (serial):', times(3), s.last_timing_ms())
s.enable_timing(False)
print('then concurrent                    :', times(8))
s.enable_timing(True); print('serial again', times(2)); s.enable_timing(False)
print('concurrent again                   :', times(8))
PY
