"""Timed CPU baseline ("cvxopt-class stand-in") for bench.py.  TEST/BASELINE INFRASTRUCTURE ONLY.

What one CPU "solve" is: the reference's dense assembly (oracle restatement of MPC.py:187-286,
prints removed), a dense primal-dual interior-point solve on the full 25h-variable problem stopped
at cvxopt's default tolerances (``qp_exact.solve_ipm_fullsize`` - the real cvxopt is not installable
here), and ``lowLevelControl`` (MPC.py:444-470).  One process per host core, BLAS pinned to one
thread per process (8 BLAS threads made a single solve 13x slower in the survey probe).
"""
from __future__ import annotations

import os
import time

import numpy as np


_BLAS_LIMIT = None  # must stay referenced: the limit is undone when the object is collected


def _init_worker():
    global _BLAS_LIMIT
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _BLAS_LIMIT = threadpool_limits(limits=1)
    except Exception:
        pass


def _solve_chunk(args):
    seed, count, exact = args
    from oracle import qp_exact
    from oracle import reference_mpc as rm
    mpc, biped = rm.MPCParams(), rm.BipedParams()
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    t_asm = 0.0
    for _ in range(count):
        x_fb = np.concatenate([rng.uniform(-0.3, 0.3, 3), rng.uniform(-1, 1, 2), rng.uniform(0.45, 0.60, 1),
                               rng.normal(0, 0.5, 3), rng.normal(0, 0.3, 3)])
        q = rm.Q0 + rng.normal(0, 0.1, 10)
        qd = rng.normal(0, 0.5, 10)
        t = float(rng.uniform(0, 0.8))
        gait = int(rng.uniform() < 0.85)
        pf_w = rm.getFootPositionWorld(x_fb, q, biped)
        contact = rm.get_contact_sequence(t, mpc) if gait else np.ones((mpc.h, 2))
        ta = time.perf_counter()
        qp = rm.build_qp(x_fb, t, pf_w.reshape(-1), mpc, biped, contact)
        t_asm += time.perf_counter() - ta
        if exact:
            z = qp_exact.solve(qp["H"], qp["f"], qp["G"], qp["hv"], qp["A"], qp["b"])["x"]
        else:
            z, _ = qp_exact.solve_ipm_fullsize(qp["H"], qp["f"], qp["G"], qp["hv"], qp["A"], qp["b"])
        u0 = z[13 * mpc.h:13 * mpc.h + 12].reshape(-1, 1)
        rm.lowLevelControl(x_fb, t, pf_w, q, qd, mpc, biped, contact, u0)
    return count, time.perf_counter() - t0, t_asm


def run(per_core: int = 12, cores: int | None = None, exact: bool = False, seed: int = 20250106):
    """Run ``per_core`` solves on each of ``cores`` processes; return a dict for bench.py."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores, initializer=_init_worker) as pool:
        res = pool.map(_solve_chunk, [(seed + 1000 + i, per_core, exact) for i in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    per_solve = float(np.mean([r[1] / r[0] for r in res]))
    asm = float(np.mean([r[2] / r[0] for r in res]))
    return dict(solves=total, cores=cores, wall_s=wall, busy_s=busy, solves_per_s=total / busy,
                per_core_solves_per_s=1.0 / per_solve, ms_per_solve=1e3 * per_solve, ms_assembly=1e3 * asm)


if __name__ == "__main__":
    print(run())
    print(run(exact=True))
