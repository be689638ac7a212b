"""Synthetic randomised biped states (SURVEY.md 8d / BASELINE.md 4) for parity tests and bench.py.

``rng = numpy.random.default_rng(20250106 + shard_index)``; euler ~ U(-0.3,0.3) rad; x,y ~ U(-1,1) m;
z ~ U(0.45,0.60) m; omega ~ N(0,0.5^2); v ~ N(0,0.3^2); q = nominal (MPC.py:15) + N(0,0.1^2);
qd ~ N(0,0.5^2); t ~ U(0,0.8) s; walking with probability 0.85 else standing;
foot = pf_w = getFootPositionWorld(x_fb, q, biped) (MPC.py:406-424) so feet agree with the joints.

The forward kinematics here is a vectorised host-side restatement used only to *generate inputs*
(the product's own FK is the CUDA kernel behind ``BatchedMPC.foot_positions``).
"""
from __future__ import annotations

import numpy as np

from .gait import batch_contact_and_phase
from .params import MPC, Biped

SEED = 20250106
Q_NOMINAL = np.array([0, 0, -np.pi / 4, np.pi / 2, -np.pi / 4, 0, 0, -np.pi / 4, np.pi / 2, -np.pi / 4])


def _eul2rotm_batch(e):
    cr, cp, cy = np.cos(e[:, 0]), np.cos(e[:, 1]), np.cos(e[:, 2])
    sr, sp, sy = np.sin(e[:, 0]), np.sin(e[:, 1]), np.sin(e[:, 2])
    R = np.empty((e.shape[0], 3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = -sp, cp * sr, cp * cr
    return R


def _foot_body_batch(q, side):
    s0, c0, s1, c1 = np.sin(q[:, 0]), np.cos(q[:, 0]), np.sin(q[:, 1]), np.cos(q[:, 1])
    s2, c2, s3, c3 = np.sin(q[:, 2]), np.cos(q[:, 2]), np.sin(q[:, 3]), np.cos(q[:, 3])
    s4, c4 = np.sin(q[:, 4]), np.cos(q[:, 4])
    fc, fs = c0 * c2 - s0 * s1 * s2, c0 * s2 + c2 * s0 * s1
    lc, ls = c2 * s0 + c0 * s1 * s2, s0 * s2 - c0 * c2 * s1
    px = (-(3 * c0) / 200 - (9 * s4 * (c3 * fc - s3 * fs)) / 250 - (11 * c0 * s2) / 50 - (side * s0) / 50
          - (11 * c3 * fs) / 50 - (11 * s3 * fc) / 50 - (9 * c4 * (c3 * fs + s3 * fc)) / 250
          - (23 * c1 * side * s0) / 1000 - (11 * c2 * s0 * s1) / 50)
    py = ((c0 * side) / 50 - (9 * s4 * (c3 * lc - s3 * ls)) / 250 - (3 * s0) / 200 - (11 * s0 * s2) / 50
          - (11 * c3 * ls) / 50 - (11 * s3 * lc) / 50 - (9 * c4 * (c3 * ls + s3 * lc)) / 250
          + (23 * c0 * c1 * side) / 1000 + (11 * c0 * c2 * s1) / 50)
    pz = ((23 * side * s1) / 1000 - (11 * c1 * c2) / 50 - (9 * c4 * (c1 * c2 * c3 - c1 * s2 * s3)) / 250
          + (9 * s4 * (c1 * c2 * s3 + c1 * c3 * s2)) / 250 - (11 * c1 * c2 * c3) / 50 + (11 * c1 * s2 * s3) / 50
          - 3.0 / 50.0)
    return np.stack([px, py, pz], axis=1)


def foot_positions_world(x_fb, q, biped):
    """Vectorised getFootPositionWorld (MPC.py:406-424): (N,12),(N,10) -> (N,6)."""
    R = _eul2rotm_batch(x_fb[:, 0:3])
    hip = np.asarray(biped.hip_offset, dtype=np.float64)
    out = np.empty((x_fb.shape[0], 6))
    for leg, side in enumerate((1.0, -1.0)):
        pb = _foot_body_batch(q[:, 5 * leg:5 * leg + 5], side) + np.array([hip[0], side * hip[1], hip[2]])
        out[:, 3 * leg:3 * leg + 3] = x_fb[:, 3:6] + np.einsum("nji,nj->ni", R, pb)  # R.T @ pb
    return out


def make_batch(n: int, shard_index: int = 0, mpc=None, biped=None, walking_prob: float = 0.85,
               extend: bool = False):
    """N synthetic instances -> dict of contiguous numpy arrays ready for ``BatchedMPC.step_host``."""
    mpc = mpc if mpc is not None else MPC()
    biped = biped if biped is not None else Biped()
    rng = np.random.default_rng(SEED + int(shard_index))
    x_fb = np.concatenate([rng.uniform(-0.3, 0.3, (n, 3)), rng.uniform(-1, 1, (n, 2)), rng.uniform(0.45, 0.60, (n, 1)),
                           rng.normal(0, 0.5, (n, 3)), rng.normal(0, 0.3, (n, 3))], axis=1)
    q = Q_NOMINAL[None, :] + rng.normal(0, 0.1, (n, 10))
    qd = rng.normal(0, 0.5, (n, 10))
    t = rng.uniform(0, 0.8, n)
    gait = (rng.uniform(size=n) < walking_prob).astype(np.int32)
    pf_w = foot_positions_world(x_fb, q, biped)
    contact, phase_k = batch_contact_and_phase(t, gait, mpc, extend=extend)
    return dict(x_fb=np.ascontiguousarray(x_fb), t=t, q=q, qd=qd, gait=gait, pf_w=pf_w, foot=pf_w.copy(),
                contact=contact, phase_k=phase_k)


def rollout_biped():
    """Robot parameters of the closed-loop rollout: the reference's ``Biped`` (MPC.py:34-48) with a
    symmetric horizontal force box ``f_min = [-f_max_x, -f_max_y, 0]``.

    With the reference default ``f_min = 0`` (MPC.py:46) the box rows force fx, fy >= 0 (MPC.py:245-246),
    so a robot can never brake a positive velocity: its closed loop drifts and diverges within ~150 ticks
    even from the nominal standing state (measured with oracle/rollout.py).  The single-tick parity
    targets all use the reference defaults; only the rollout workload uses this variant."""
    b = Biped()
    b.f_min = np.array([[-float(b.f_max[0, 0])], [-float(b.f_max[1, 0])], [0.0]])
    return b


def make_rollout_batch(n: int, shard_index: int = 0, biped=None, walking_prob: float = 0.85):
    """Initial conditions of the closed-loop rollout (BASELINE.json configs[4]): the reference's nominal
    state (MPC.py:13-16) plus a small perturbation - euler ~ U(-0.05,0.05) rad, x,y,z offsets ~ U(-0.02,0.02) m,
    omega, v ~ N(0,0.05^2), q = nominal + N(0,0.02^2) - a random integer gait clock in [0,10) and
    ``foot = getFootPositionWorld(x, q)``; walking with probability 0.85 else standing."""
    biped = biped if biped is not None else rollout_biped()
    rng = np.random.default_rng(SEED + 7919 + int(shard_index))
    x = np.array([0, 0, 0, 0, 0, 0.53, 0, 0, 0, 0, 0, 0], dtype=np.float64)[None, :] + np.concatenate(
        [rng.uniform(-0.05, 0.05, (n, 3)), rng.uniform(-0.02, 0.02, (n, 3)), rng.normal(0, 0.05, (n, 3)),
         rng.normal(0, 0.05, (n, 3))], axis=1)
    q = Q_NOMINAL[None, :] + rng.normal(0, 0.02, (n, 10))
    qd = np.zeros((n, 10))
    tick = rng.integers(0, 10, n).astype(np.int32)
    gait = (rng.uniform(size=n) < walking_prob).astype(np.uint8)
    foot = foot_positions_world(x, q, biped)
    return dict(x=np.ascontiguousarray(x), foot=np.ascontiguousarray(foot), tick=tick, gait=gait, q=q, qd=qd)
