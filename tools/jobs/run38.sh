cd $GRAFT_REPO_ROOT
for n in 32768 65536 131072 262144; do
EXP=both python - $n <<'PY' 2>&1 | tail -2
import sys, os, numpy as np, torch
sys.path.insert(0, '.')
import biped_mpc_py_b200._lib as _l
_l.LIB_PATH = os.path.join(os.path.dirname(_l.LIB_PATH), "_exp", "lib_both.so")
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = int(sys.argv[1])
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
    return ts, out
for both in (1, 0):
    s.set_option("lane_both", both)
    times(3)
    ts, out = times(8)
    st = np.bincount(out["status"].cpu().numpy(), minlength=4).tolist()
    ref = out["controls"].clone() if both else ref
    print(f"n={n} lane_both={both}: median {np.median(ts):.2f} ms min {min(ts):.2f} = {n/np.median(ts)/1e3:.2f} M solves/s status {st}" + ("" if both else f" max abs diff vs both {float((out['controls']-ref).abs().max()):.1e}"), flush=True)
PY
done
