cd $GRAFT_REPO_ROOT
python - <<'PY' 2>&1 | tail -12
import sys, numpy as np, torch, time
sys.path.insert(0, '.')
from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
n = 65536
mpc = MPC(h=30)
s = BatchedMPC(mpc, Biped(), max_batch=n, extend_gait=True)
for shard in range(1000, 1008):
    b = synth.make_batch(n, shard_index=shard, mpc=mpc, extend=True)
    t0 = time.time()
    out = s.step_host(b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"], phase_k=b["phase_k"])
    dt = time.time() - t0
    bad = np.nonzero(out["status"] != 0)[0]
    print(f"shard {shard}: {dt*1e3:.0f} ms wall, status {np.bincount(out['status'], minlength=4).tolist()} bad idx {bad.tolist()} gait {b['gait'][bad].tolist()} iters {out['iters'][bad].tolist()}", flush=True)
    if len(bad):
        np.savez('gpurun_out/h30_bad_%d.npz' % shard, idx=bad, **{k: b[k][bad] for k in ("x_fb","t","foot","contact","q","qd","pf_w","phase_k","gait")})
PY
