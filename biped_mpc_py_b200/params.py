"""Parameter objects with the reference's attribute names, and their C-ABI packing.

``MPC`` / ``Biped`` mirror the attribute sets of the reference classes (MPC.py:22-48); any
object exposing the same attributes (e.g. the reference's own instances) is accepted by the
solver, which reads them duck-typed exactly as the reference functions do.
"""
from __future__ import annotations

import ctypes

import numpy as np


class MPC:
    """Horizon, weights and swing-leg gains (attribute names of MPC.py:22-32)."""

    def __init__(self, h: int = 10):
        self.h = h
        self.dt = 0.04
        self.x_cmd = np.array([0, 0, 0, 0, 0, 0.55, 0, 0, 0, 0, 0, 0], dtype=np.float64)
        self.Q = np.array([500, 100, 100, 300, 300, 700, 1, 1, 1, 1, 1, 1, 1], dtype=np.float64)
        self.R = np.ones(12) * 1e-4
        self.kv = 0.01
        self.kp = 500.0 * np.eye(3)
        self.kd = 10.0 * np.eye(3)
        self.swingHeight = 0.1


class Biped:
    """Robot constants and force/moment limits (attribute names of MPC.py:34-48)."""

    def __init__(self):
        self.m = 12
        self.I = np.diag([0.932, 0.9420, 0.0711])
        self.lt = 0.09
        self.lh = 0.05
        self.g = 9.81
        self.hip_offset = np.array([-0.005, 0.047, -0.126])
        self.mu = 0.5
        self.f_max = np.full((3, 1), 500.0)
        self.f_min = np.zeros((3, 1))
        self.tau_max = np.array([[0.0], [67.0], [33.5]])
        self.tau_min = -self.tau_max


class BmpcParams(ctypes.Structure):
    """``struct bmpc_params`` of include/biped_mpc_b200.h (field order and types must match)."""

    _fields_ = [
        ("h", ctypes.c_int32), ("extend_gait", ctypes.c_int32), ("dt", ctypes.c_double),
        ("x_cmd", ctypes.c_double * 12), ("Q", ctypes.c_double * 13), ("R", ctypes.c_double * 12),
        ("kv", ctypes.c_double), ("kp", ctypes.c_double * 9), ("kd", ctypes.c_double * 9),
        ("swing_height", ctypes.c_double), ("mass", ctypes.c_double), ("inertia", ctypes.c_double * 9),
        ("lt", ctypes.c_double), ("lh", ctypes.c_double), ("g", ctypes.c_double),
        ("hip_offset", ctypes.c_double * 3), ("mu", ctypes.c_double),
        ("f_max", ctypes.c_double * 3), ("f_min", ctypes.c_double * 3),
        ("tau_max", ctypes.c_double * 3), ("tau_min", ctypes.c_double * 3),
        ("max_iter", ctypes.c_int32), ("polish", ctypes.c_int32),
        ("mu_tol", ctypes.c_double), ("rd_tol", ctypes.c_double),
    ]


def _vec(dst, src, n):
    a = np.asarray(src, dtype=np.float64).reshape(-1)
    if a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    for i in range(n):
        dst[i] = float(a[i])


def pack_params(mpc, biped, extend_gait: bool = False, max_iter: int = 0, mu_tol: float = 0.0,
                rd_tol: float = 0.0) -> BmpcParams:
    """Read ``mpc`` / ``biped`` by attribute name (as MPC.py does) into the C struct."""
    p = BmpcParams()
    p.h = int(mpc.h)
    p.extend_gait = int(bool(extend_gait))
    p.dt = float(mpc.dt)
    _vec(p.x_cmd, mpc.x_cmd, 12)
    _vec(p.Q, mpc.Q, 13)
    _vec(p.R, mpc.R, 12)
    p.kv = float(mpc.kv)
    _vec(p.kp, mpc.kp, 9)
    _vec(p.kd, mpc.kd, 9)
    p.swing_height = float(mpc.swingHeight)
    p.mass = float(biped.m)
    _vec(p.inertia, biped.I, 9)
    p.lt, p.lh, p.g = float(biped.lt), float(biped.lh), float(biped.g)
    _vec(p.hip_offset, biped.hip_offset, 3)
    p.mu = float(biped.mu)
    _vec(p.f_max, biped.f_max, 3)
    _vec(p.f_min, biped.f_min, 3)
    _vec(p.tau_max, biped.tau_max, 3)
    _vec(p.tau_min, biped.tau_min, 3)
    p.max_iter, p.polish, p.mu_tol, p.rd_tol = int(max_iter), 0, float(mu_tol), float(rd_tol)
    return p


def params_key(p: BmpcParams) -> bytes:
    return bytes(p)
