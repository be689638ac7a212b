"""Oracle solves spread over the host cores (spawned workers, one BLAS thread each) for the large parity tests."""
import os

import numpy as np


def _init():
    for v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"


def _chunk(args):
    batch, idx, h, extend = args
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import reference_mpc as rm
    mpc, biped = rm.MPCParams(h=h), rm.BipedParams()
    U, T, M = [], [], []
    for i in idx:
        states, controls = rm.solve_mpc(batch["x_fb"][i], float(batch["t"][i]), batch["foot"][i], mpc, biped,
                                        batch["contact"][i], extend=extend)
        tau = rm.lowLevelControl(batch["x_fb"][i], float(batch["t"][i]), batch["pf_w"][i].reshape(6, 1), batch["q"][i],
                                 batch["qd"][i], mpc, biped, batch["contact"][i], controls[0].reshape(-1, 1)).reshape(-1)
        scale = max(1.0, np.abs(controls).max())
        M.append([rm.active_friction_rows(controls[s], batch["contact"][i][s], biped.mu, scale) for s in range(h)])
        U.append(controls), T.append(tau)
    return idx, np.array(U), np.array(T), np.array(M, dtype=np.uint8)


def oracle_parallel(batch, n, h=10, extend=False, workers=None):
    """Returns controls (n,h,12), tau (n,10), friction masks (n,h) of the oracle for instances 0..n-1."""
    import multiprocessing as mp
    workers = workers or max(1, len(os.sched_getaffinity(0)))
    keys = ("x_fb", "t", "foot", "contact", "q", "qd", "pf_w")
    small = {k: np.asarray(batch[k])[:n] for k in keys}
    chunks = [list(range(s, min(n, s + 32))) for s in range(0, n, 32)]
    U = np.zeros((n, h, 12)); T = np.zeros((n, 10)); M = np.zeros((n, h), dtype=np.uint8)
    saved = {v: os.environ.get(v) for v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS")}
    _init()  # spawned workers inherit the environment: one BLAS thread each (8 threads make one solve 13x slower)
    try:
        with mp.get_context("spawn").Pool(workers, initializer=_init) as pool:
            for idx, u, t, m in pool.imap_unordered(_chunk, [(small, c, h, extend) for c in chunks]):
                U[idx], T[idx], M[idx] = u, t, m
    finally:
        for v, val in saved.items():
            if val is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = val
    return U, T, M
