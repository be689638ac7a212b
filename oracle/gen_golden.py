"""Generate tests/golden/reference_cases.npz by running the REAL reference module.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no
/root/reference):

    python -m oracle.gen_golden

The reference (``/root/reference/bipedalLocomotionMPC.py``) imports ``cvxopt`` at
MPC.py:3 and executes its driver script at import (MPC.py:475-495).  cvxopt is not
installable here, so a stub module is injected that (a) records the dense
matrices the reference hands to ``cvxopt.solvers.qp`` (MPC.py:289-297) and (b)
answers with the certified exact optimum from ``oracle.qp_exact``.  Everything
else - references, dynamics, constraint rows, cost, Jacobians, swing PD, forward
kinematics, gait table - is computed by the reference's own, unmodified code.

What the fixture pins:
  * the reference's own assembly (H, f, G, h, A, b) for several cases, dense;
  * its helper outputs (contact table sweep, x_ref, foot_ref, A_k/B_k, Jacobians,
    forward kinematics, swing force, torques) for 64 cases and 3 parameter variants;
  * states/controls = exact optimum of the reference-built QP, and the torques the
    reference's ``lowLevelControl`` derives from them.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

REFERENCE_FILE = "/root/reference/bipedalLocomotionMPC.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_cases.npz")
SEED = 20250106
N_CASES = 64
N_DENSE = 6


def load_reference():
    from oracle import qp_exact

    captured = {}
    stub = types.ModuleType("cvxopt")
    stub.matrix = lambda a: np.array(a, dtype=float)
    stub.solvers = types.SimpleNamespace()

    def qp(H, f, G=None, h=None, A=None, b=None):
        captured.update(H=H, f=f, G=G, hv=np.reshape(h, -1), A=A, b=b)
        sol = qp_exact.solve(H, f, G, np.reshape(h, -1), A, b)
        captured["sol"] = sol
        return {"x": sol["x"].reshape(-1, 1), "status": "optimal"}

    stub.solvers.qp = qp
    sys.modules["cvxopt"] = stub
    spec = importlib.util.spec_from_file_location("reference_mpc_module", REFERENCE_FILE)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod, captured


def variant_params(ref, variant):
    mpc, biped = ref.MPC(), ref.Biped()
    if variant == 1:  # commanded velocities exercise the "!= 0" branch of MPC.py:66-67
        mpc.x_cmd = np.array([0, 0, 0, 0, 0, 0.55, 0.2, 0, 0, 0.3, -0.1, 0], dtype=float)
    elif variant == 2:  # different limits: mx unpinned, negative tangential minimum, other mu/weights
        biped.mu = 0.7
        biped.f_min = np.array([[-120.0], [-80.0], [0.0]])
        biped.f_max = np.array([[300.0], [300.0], [400.0]])
        biped.tau_max = np.array([[4.0], [50.0], [20.0]])
        biped.tau_min = np.array([[-3.0], [-45.0], [-20.0]])
        biped.m = 13.5
        biped.I = np.array([[0.9, 0.02, 0.01], [0.02, 0.95, -0.015], [0.01, -0.015, 0.08]])
        mpc.Q = np.array([400, 150, 120, 250, 320, 650, 2, 1.5, 1, 1, 2, 1, 1], dtype=float)
        mpc.R = np.array([1, 2, 1, 1, 2, 1, 3, 1, 2, 3, 1, 2], dtype=float) * 1e-4
        mpc.kv = 0.02
    return mpc, biped


def main():
    ref, captured = load_reference()
    rng = np.random.default_rng(SEED)
    quiet = contextlib.redirect_stdout(io.StringIO())

    keys = ["x_fb", "t", "q", "qd", "gait", "variant", "pf_w", "contact", "x_ref", "foot_ref", "A0", "B0", "A1",
            "B1", "states", "controls", "tau", "obj", "Jm", "vf_w", "F_swing", "cert_stationarity"]
    out = {k: [] for k in keys}
    dense = {k: [] for k in ["H", "f", "G", "hv", "A", "b"]}

    q_nom = np.array([0, 0, -np.pi / 4, np.pi / 2, -np.pi / 4, 0, 0, -np.pi / 4, np.pi / 2, -np.pi / 4])
    for c in range(N_CASES):
        variant = 0 if c < 40 else (1 if c < 52 else 2)
        mpc, biped = variant_params(ref, variant)
        if c == 0:  # the reference script's own state (G1)
            x_fb, q, qd, t, gait = np.array([0, 0, 0, 0, 0, 0.53, 0, 0, 0, 0, 0, 0.0]), q_nom.copy(), np.zeros(10), 0.0, 1
        elif c == 1:  # G4 standing
            x_fb, q, qd, t, gait = np.array([0, 0, 0, 0, 0, 0.53, 0, 0, 0, 0, 0, 0.0]), q_nom.copy(), np.zeros(10), 0.0, 0
        else:  # SURVEY.md 8d distribution
            x_fb = np.concatenate([rng.uniform(-0.3, 0.3, 3), rng.uniform(-1, 1, 2), rng.uniform(0.45, 0.60, 1),
                                   rng.normal(0, 0.5, 3), rng.normal(0, 0.3, 3)])
            q = q_nom + rng.normal(0, 0.1, 10)
            qd = rng.normal(0, 0.5, 10)
            t = float(rng.uniform(0, 0.8))
            gait = int(rng.uniform() < 0.85)
        with quiet:
            pf_w = ref.getFootPositionWorld(x_fb, q, biped)
            foot = pf_w.reshape(-1)
            contact = ref.get_contact_sequence(t, mpc) if gait == 1 else np.ones((mpc.h, 2))
            x_ref = ref.get_reference_trajectory(x_fb, mpc)
            foot_ref = ref.get_reference_foot_trajectory(x_fb, t, foot, mpc, contact)
            A0, B0 = ref.get_simplified_dynamics(mpc, biped, x_ref[:, 0], foot_ref[:, 0])
            A1, B1 = ref.get_simplified_dynamics(mpc, biped, x_ref[:, 7], foot_ref[:, 7])
            states, controls = ref.solve_mpc(x_fb, t, foot, mpc, biped, contact)
            u0 = controls[0, :].reshape(-1, 1)
            tau = ref.lowLevelControl(x_fb, t, pf_w, q, qd, mpc, biped, contact, u0)
            jm = np.stack([ref.getLegKinematics(*q[5 * leg:5 * leg + 5], side)[0] for leg, side in ((0, 1), (1, -1))])
            rot = ref.eul2rotm(x_fb[0:3])
            vf = np.stack([(rot.T @ jm[leg][0:3] @ qd[5 * leg:5 * leg + 5].reshape(-1, 1)) for leg in range(2)])
            fs = np.stack([ref.swingLegControl(x_fb, t, pf_w[3 * leg:3 * leg + 3], vf[leg], mpc, side)
                           for leg, side in ((0, 1), (1, -1))])
        vals = dict(x_fb=x_fb, t=t, q=q, qd=qd, gait=gait, variant=variant, pf_w=pf_w.reshape(-1),
                    contact=np.asarray(contact, dtype=np.uint8), x_ref=x_ref, foot_ref=foot_ref, A0=A0, B0=B0, A1=A1,
                    B1=B1, states=states, controls=controls, tau=tau.reshape(-1), obj=captured["sol"]["obj"], Jm=jm,
                    vf_w=vf.reshape(2, 3), F_swing=fs.reshape(2, 3),
                    cert_stationarity=captured["sol"]["cert"]["stationarity"])
        for k in keys:
            out[k].append(vals[k])
        if c < N_DENSE or c in (40, 52):
            for k in dense:
                dense[k].append(np.array(captured[k], dtype=float))
    arrays = {k: np.array(v) for k, v in out.items()}
    arrays["dense_case_index"] = np.array(list(range(N_DENSE)) + [40, 52])
    for k, v in dense.items():
        arrays["dense_" + k] = np.array(v)

    # gait table / float phase sweep (MPC.py:50-59): every 5 ms over two gait cycles and then some
    mpc = ref.MPC()
    ts = np.round(np.arange(0.0, 1.3, 0.005), 6)
    ts = np.concatenate([ts, np.arange(0, 61) * 0.04])  # exact tick multiples hit the float floor-division quirk
    arrays["sweep_t"] = ts
    arrays["sweep_phase"] = np.array([int(t // mpc.dt) for t in ts])
    arrays["sweep_contact"] = np.array([ref.get_contact_sequence(t, mpc) for t in ts], dtype=np.uint8)

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", N_CASES, "cases; max stationarity",
          float(arrays["cert_stationarity"].max()))


if __name__ == "__main__":
    main()
