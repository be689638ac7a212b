cd $GRAFT_REPO_ROOT
for n in 65536 262144; do
BMPC_TRACE=1 python - $n <<'PY' 2>&1 | grep "trace\|tick" | tail -7
import sys, os, numpy as np, torch
sys.path.insert(0, '.')
import biped_mpc_py_b200._lib as _l
_l.LIB_PATH = os.path.join(os.path.dirname(_l.LIB_PATH), "_exp", "lib_trace.so")
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = int(sys.argv[1])
mpc, biped = MPC(), Biped()
b = synth.make_batch(n, shard_index=0, mpc=mpc, biped=biped)
s = BatchedMPC(mpc, biped, max_batch=n)
dev = s.device
tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8), tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = s.step(*d); e1.record(); torch.cuda.synchronize()
    print('tick', n, round(e0.elapsed_time(e1), 2), flush=True)
PY
done
