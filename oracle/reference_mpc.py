"""Float64 restatement of the reference's MPC formulation and low-level control.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Citations ``MPC.py:a-b`` are
into /root/reference/bipedalLocomotionMPC.py.  Written from the reference's
behaviour, in this repo's own structure; the quirks listed in SURVEY.md 8a are kept
on purpose because the product must match the reference, not fix it.

State  x  = [euler(3), position(3), omega(3), velocity(3), 1]      (MPC.py:9)
Input  u  = [f1(3), f2(3), m1(3), m2(3)]                           (MPC.py:10)
Decision  z = [X_0..X_{h-1}, U_0..U_{h-1}],  X_i = state after U_i (MPC.py:209-213)
"""
from __future__ import annotations

import numpy as np

from . import qp_exact

NX, NU = 13, 12


class MPCParams:
    """Mirror of ``class MPC`` (MPC.py:22-32)."""

    def __init__(self, h: int = 10):
        self.h = h
        self.dt = 0.04
        self.x_cmd = np.array([0, 0, 0, 0, 0, 0.55, 0, 0, 0, 0, 0, 0], dtype=float)
        self.Q = np.array([500, 100, 100, 300, 300, 700, 1, 1, 1, 1, 1, 1, 1], dtype=float)
        self.R = np.full(12, 1e-4)
        self.kv = 0.01
        self.kp = np.eye(3) * 500.0
        self.kd = np.eye(3) * 10.0
        self.swingHeight = 0.1


class BipedParams:
    """Mirror of ``class Biped`` (MPC.py:34-48)."""

    def __init__(self):
        self.m = 12
        self.I = np.diag([0.932, 0.9420, 0.0711])
        self.lt = 0.09
        self.lh = 0.05
        self.g = 9.81
        self.hip_offset = np.array([-0.005, 0.047, -0.126])
        self.mu = 0.5
        self.f_max = np.array([[500.0], [500.0], [500.0]])
        self.f_min = np.array([[0.0], [0.0], [0.0]])
        self.tau_max = np.array([[0.0], [67.0], [33.5]])
        self.tau_min = -self.tau_max


# Default joint configuration and state of the reference script (MPC.py:13-16).
X_FB0 = np.array([0, 0, 0, 0, 0, 0.53, 0, 0, 0, 0, 0, 0], dtype=float)
Q0 = np.array([0, 0, -np.pi / 4, np.pi / 2, -np.pi / 4, 0, 0, -np.pi / 4, np.pi / 2, -np.pi / 4])
QD0 = np.zeros(10)

GAIT_PERIOD = 10  # ticks; 5 left-stance then 5 right-stance (MPC.py:52-55)


def gait_phase(t, mpc) -> int:
    """``int(t // dt)`` in floating point, exactly as MPC.py:56 / MPC.py:99.

    The float floor-division is off by one at many tick boundaries (e.g.
    ``(3*0.04)//0.04 == 2.0``); that is reference behaviour and is kept.
    """
    return int(t // mpc.dt)


def get_contact_sequence(t, mpc, extend: bool = False):
    """Contact schedule, MPC.py:50-59.

    Reference: rows ``k:k+10`` of a 20x2 table with ``k = phase % h``.  With
    ``extend=True`` (new behaviour, SURVEY.md 8a "h=30 extension") the table is
    continued periodically so any horizon gets ``h`` rows; identical for h=10.
    """
    stance_left = (np.arange(2 * GAIT_PERIOD) % GAIT_PERIOD) < GAIT_PERIOD // 2
    table = np.stack([stance_left, ~stance_left], axis=1).astype(int)
    phase = gait_phase(t, mpc)
    if not extend:
        k = phase % mpc.h
        return table[k:k + 10, :]
    k = phase % GAIT_PERIOD
    rows = (k + np.arange(mpc.h)) % GAIT_PERIOD
    return table[rows, :]


def get_reference_trajectory(x_fb, mpc):
    """13 x h state reference, MPC.py:61-70."""
    h = mpc.h
    x_ref = np.empty((NX, h))
    x_ref[:12, :] = np.asarray(mpc.x_cmd, dtype=float)[:, None]
    x_ref[12, :] = 1.0
    x_ref[:12, 0] = x_fb
    for i in range(6):
        rate = mpc.x_cmd[i + 6]
        if rate != 0:
            for k in range(1, h):
                x_ref[i, k] = x_fb[i] + rate * (k * mpc.dt)
    return x_ref


def next_footholds(x_fb, mpc):
    """The two candidate footholds of MPC.py:73-93 (each sets BOTH feet to one x,y).

    Keeps the ``x_fb[10]`` start of ``foot_des_y_2`` (MPC.py:87) and the 0.1*v vs
    0.2*v look-ahead (``1/2*h/2*dt`` vs ``1/2*h*dt``).
    """
    h, dt, kv = mpc.h, mpc.dt, mpc.kv
    ex = kv * (x_fb[3] - mpc.x_cmd[3])
    ey = kv * (x_fb[4] - mpc.x_cmd[4])
    x1 = x_fb[3] + x_fb[9] * 1 / 2 * h / 2 * dt + ex
    x2 = x_fb[3] + x_fb[9] * 1 / 2 * h * dt + ex
    y1 = x_fb[4] + x_fb[10] * 1 / 2 * h / 2 * dt + ey
    y2 = x_fb[10] + x_fb[10] * 1 / 2 * h * dt + ey
    return np.array([x1, y1, 0.0, x1, y1, 0.0]), np.array([x2, y2, 0.0, x2, y2, 0.0])


def get_reference_foot_trajectory(x_fb, t, foot, mpc, contact, extend: bool = False):
    """6 x h foot-position reference, MPC.py:72-109.

    Single support at stage 0: ``5-kk`` columns of the current feet, 5 of
    foothold 1, ``kk`` of foothold 2 (always 10 columns in the reference,
    MPC.py:101-106).  ``extend=True`` pads with foothold 2 / truncates to ``h``.
    """
    foot = np.asarray(foot, dtype=float).reshape(6)
    f1, f2 = next_footholds(x_fb, mpc)
    k = gait_phase(t, mpc) % (GAIT_PERIOD if extend else mpc.h)
    kk = k % 5
    if np.sum(contact[0, :]) == 1:
        cols = [foot] * (5 - kk) + [f1] * 5
        if extend:
            cols = (cols + [f2] * max(mpc.h - len(cols), 0))[:mpc.h]
        else:
            cols = cols + [f2] * kk  # always 10 columns, whatever mpc.h is
        return np.stack(cols, axis=1)
    return np.tile(foot[:, None], (1, mpc.h))


def eul2rotm(eul):
    """Rz(eul[2]) @ Ry(eul[1]) @ Rx(eul[0]), MPC.py:111-138 (eul = [roll,pitch,yaw])."""
    (cr, cp, cy), (sr, sp, sy) = np.cos(eul), np.sin(eul)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])


def skew(v):
    """MPC.py:140-146."""
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def dynamics_rotation(yaw, pitch, roll):
    """scipy ``Rotation.from_euler('zyx', [yaw, pitch, roll])`` (MPC.py:154-156).

    Lower-case axes are *extrinsic*: rotate about z, then y, then x of the fixed
    frame, i.e. the matrix Rx(roll) @ Ry(pitch) @ Rz(yaw).  Checked against scipy in
    tests/test_oracle_units.py.
    """
    cz, sz, cy, sy, cx, sx = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    ry = np.array([[cy, 0, sy], [0, 1.0, 0], [-sy, 0, cy]])
    rx = np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return rx @ ry @ rz


def euler_rate_map(yaw, pitch):
    """Inverse of the 3x3 at MPC.py:160-164 (maps world omega to euler rates)."""
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    return np.linalg.inv(np.array([[cy * cp, -sy, 0.0], [sy * cp, cy, 0.0], [-sp, 0.0, 1.0]]))


def get_simplified_dynamics(mpc, biped, x_ref_k, foot_ref_k):
    """Forward-Euler single-rigid-body A (13x13), B (13x12), MPC.py:148-185.

    NOTE the dynamics read x[0] as yaw and x[2] as roll (MPC.py:151-153) while
    ``eul2rotm`` reads x[0] as roll - reference quirk, kept.
    """
    yaw, pitch, roll = x_ref_k[0], x_ref_k[1], x_ref_k[2]
    rot = dynamics_rotation(yaw, pitch, roll)
    inertia_w = rot.T @ np.asarray(biped.I, dtype=float) @ rot
    a_c = np.zeros((NX, NX))
    a_c[0:3, 6:9] = euler_rate_map(yaw, pitch)
    a_c[3:6, 9:12] = np.eye(3)
    a_c[11, 12] = -biped.g
    b_c = np.zeros((NX, NU))
    com = x_ref_k[3:6]
    for leg in range(2):
        lever = foot_ref_k[3 * leg:3 * leg + 3] - com
        b_c[6:9, 3 * leg:3 * leg + 3] = np.linalg.solve(inertia_w, skew(lever))
        b_c[6:9, 6 + 3 * leg:9 + 3 * leg] = np.linalg.solve(inertia_w, np.eye(3))
        b_c[9:12, 3 * leg:3 * leg + 3] = np.eye(3) / biped.m
    return a_c * mpc.dt + np.eye(NX), b_c * mpc.dt


def build_qp(x_fb, t, foot, mpc, biped, contact, extend: bool = False):
    """Dense QP data exactly as MPC.py:187-286 hands to cvxopt.

    Returns dict(H, f, G, hv, A, b, x_ref, foot_ref): min 1/2 z'Hz + f'z s.t.
    G z <= hv, A z = b, with z = [X_0..X_{h-1}; U_0..U_{h-1}].
    """
    h = mpc.h
    x_fb = np.asarray(x_fb, dtype=float).reshape(12)
    contact = np.asarray(contact)
    x_ref = get_reference_trajectory(x_fb, mpc)
    foot_ref = get_reference_foot_trajectory(x_fb, t, foot, mpc, contact, extend=extend)
    rot_now = eul2rotm(x_fb[0:3])
    nxs = NX * h
    n = (NX + NU) * h

    # equality block, MPC.py:202-216:  X_i - B_i U_i - A_i X_{i-1} = 0, X_{-1} = [x_fb;1]
    a_eq = np.zeros((nxs, n))
    b_eq = np.zeros(nxs)
    for i in range(h):
        a_i, b_i = get_simplified_dynamics(mpc, biped, x_ref[:, i], foot_ref[:, i])
        rows = slice(NX * i, NX * (i + 1))
        a_eq[rows, rows] = np.eye(NX)
        a_eq[rows, nxs + NU * i:nxs + NU * (i + 1)] = -b_i
        if i == 0:
            b_eq[rows] = a_i @ np.append(x_fb, 1.0)
        else:
            a_eq[rows, NX * (i - 1):NX * i] = -a_i

    # inequality block, MPC.py:218-274 (only input columns are touched)
    mu = biped.mu
    fric_leg = np.array([[1, 0, -mu], [0, 1, -mu], [-1, 0, -mu], [0, -1, -mu]], dtype=float)
    fric = np.zeros((8, NU))
    fric[0:4, 0:3] = fric_leg
    fric[4:8, 3:6] = fric_leg
    box = np.vstack([np.eye(NU), -np.eye(NU)])
    lt = biped.lt - 0.01
    lh = biped.lh - 0.02
    ez_w = rot_now[:, 2]  # [0,0,1] @ R.T
    ey_w = rot_now[:, 1]  # [0,1,0] @ R.T
    line = np.zeros((4, NU))
    for leg in range(2):
        fcol = slice(3 * leg, 3 * leg + 3)
        mcol = slice(6 + 3 * leg, 9 + 3 * leg)
        line[2 * leg, fcol], line[2 * leg, mcol] = -lh * ez_w, ey_w
        line[2 * leg + 1, fcol], line[2 * leg + 1, mcol] = -lt * ez_w, -ey_w
    f_max = np.asarray(biped.f_max, dtype=float).reshape(3)
    f_min = np.asarray(biped.f_min, dtype=float).reshape(3)
    t_max = np.asarray(biped.tau_max, dtype=float).reshape(3)
    t_min = np.asarray(biped.tau_min, dtype=float).reshape(3)
    g_u = np.zeros((36 * h, NU * h))
    hv = np.zeros(36 * h)
    for k in range(h):
        ucols = slice(NU * k, NU * (k + 1))
        c1, c2 = float(contact[k, 0]), float(contact[k, 1])
        g_u[8 * k:8 * (k + 1), ucols] = fric
        g_u[8 * h + 24 * k:8 * h + 24 * (k + 1), ucols] = box
        hv[8 * h + 24 * k:8 * h + 24 * (k + 1)] = np.concatenate([
            c1 * f_max, c2 * f_max, c1 * t_max, c2 * t_max,
            c1 * -f_min, c2 * -f_min, c1 * -t_min, c2 * -t_min])
        g_u[32 * h + 4 * k:32 * h + 4 * (k + 1), ucols] = line
    g_full = np.hstack([np.zeros((36 * h, nxs)), g_u])

    # cost, MPC.py:277-286
    q_diag = np.tile(np.asarray(mpc.Q, dtype=float), h)
    r_diag = np.tile(np.asarray(mpc.R, dtype=float), h)
    h_mat = 2.0 * np.diag(np.concatenate([q_diag, r_diag]))
    f_vec = 2.0 * np.concatenate([-q_diag * x_ref.T.reshape(-1), np.zeros(NU * h)])
    return dict(H=h_mat, f=f_vec, G=g_full, hv=hv, A=a_eq, b=b_eq, x_ref=x_ref, foot_ref=foot_ref)


def solve_mpc(x_fb, t, foot, mpc, biped, contact, extend: bool = False, full_certificate: bool = False,
              return_info: bool = False):
    """``solve_mpc`` of MPC.py:187-304 with the exact optimum in place of cvxopt's.

    Returns ``(states (h,13), controls (h,12))`` like MPC.py:300-304.
    """
    qp = build_qp(x_fb, t, foot, mpc, biped, contact, extend=extend)
    sol = qp_exact.solve(qp["H"], qp["f"], qp["G"], qp["hv"], qp["A"], qp["b"], full_certificate=full_certificate)
    h = mpc.h
    z = sol["x"]
    states = z[:NX * h].reshape(h, NX)
    controls = z[NX * h:].reshape(h, NU)
    if return_info:
        return states, controls, sol
    return states, controls


# ----------------------------------------------------------------------------
# low-level control
# ----------------------------------------------------------------------------

def getLegKinematics(q0, q1, q2, q3, q4, side):
    """Closed-form 6x5 leg Jacobian Jm (rows 0-2 linear, 3-5 angular), MPC.py:306-365."""
    s0, c0, s1, c1 = np.sin(q0), np.cos(q0), np.sin(q1), np.cos(q1)
    a2, a23, a234 = q2, q2 + q3, q2 + q3 + q4
    # chain reach along the sagittal plane, per number of distal links included
    sin_reach = [0.04 * np.sin(a234) + 0.22 * np.sin(a23) + 0.22 * np.sin(a2),
                 0.04 * np.sin(a234) + 0.22 * np.sin(a23),
                 0.04 * np.sin(a234)]
    cos_reach = [0.04 * np.cos(a234) + 0.22 * np.cos(a23) + 0.22 * np.cos(a2),
                 0.04 * np.cos(a234) + 0.22 * np.cos(a23),
                 0.04 * np.cos(a234)]
    lat = 0.018 * side + 0.0025
    jm = np.zeros((6, 5))
    # hip yaw (MPC.py:313-322)
    across = 0.015 * side + c1 * lat - s1 * cos_reach[0]
    along = sin_reach[0] + 0.0135
    jm[0, 0] = s0 * along + c0 * across
    jm[1, 0] = s0 * across - c0 * along
    jm[5, 0] = 1.0
    # hip roll (MPC.py:324-329)
    swing = s1 * lat + c1 * cos_reach[0]
    jm[0, 1] = -s0 * swing
    jm[1, 1] = c0 * swing
    jm[2, 1] = s1 * cos_reach[0] - c1 * lat
    jm[3, 1] = c0
    jm[4, 1] = s0
    # hip pitch, knee, ankle (MPC.py:331-362): same pattern with a shrinking chain
    for col in range(3):
        sr, cr = sin_reach[col], cos_reach[col]
        jm[0, 2 + col] = s0 * s1 * sr - c0 * cr
        jm[1, 2 + col] = -s0 * cr - c0 * s1 * sr
        jm[2, 2 + col] = c1 * sr
        jm[3, 2 + col] = -c1 * s0
        jm[4, 2 + col] = c0 * c1
        jm[5, 2 + col] = s1
    return jm, jm[0:3, :]


def getFootPositionBody(q0, q1, q2, q3, q4, side):
    """Closed-form foot position in the hip frame, MPC.py:367-404.

    Uses link constants 0.22/0.036/0.015/0.02/0.023/0.06 - different from the
    Jacobian's (0.04, 0.018...) in the reference; kept.
    """
    s0, c0, s1, c1 = np.sin(q0), np.cos(q0), np.sin(q1), np.cos(q1)
    s2, c2, s3, c3, s4, c4 = np.sin(q2), np.cos(q2), np.sin(q3), np.cos(q3), np.sin(q4), np.cos(q4)
    # unit vectors of the sagittal chain expressed before the hip-yaw/roll rotation
    fwd_c = c0 * c2 - s0 * s1 * s2   # "cos-like" term in x
    fwd_s = c0 * s2 + c2 * s0 * s1   # "sin-like" term in x
    lat_c = c2 * s0 + c0 * s1 * s2
    lat_s = s0 * s2 - c0 * c2 * s1
    px = (-(3 * c0) / 200
          - (9 * s4 * (c3 * fwd_c - s3 * fwd_s)) / 250
          - (11 * c0 * s2) / 50
          - (side * s0) / 50
          - (11 * c3 * fwd_s) / 50
          - (11 * s3 * fwd_c) / 50
          - (9 * c4 * (c3 * fwd_s + s3 * fwd_c)) / 250
          - (23 * c1 * side * s0) / 1000
          - (11 * c2 * s0 * s1) / 50)
    py = ((c0 * side) / 50
          - (9 * s4 * (c3 * lat_c - s3 * lat_s)) / 250
          - (3 * s0) / 200
          - (11 * s0 * s2) / 50
          - (11 * c3 * lat_s) / 50
          - (11 * s3 * lat_c) / 50
          - (9 * c4 * (c3 * lat_s + s3 * lat_c)) / 250
          + (23 * c0 * c1 * side) / 1000
          + (11 * c0 * c2 * s1) / 50)
    pz = ((23 * side * s1) / 1000
          - (11 * c1 * c2) / 50
          - (9 * c4 * (c1 * c2 * c3 - c1 * s2 * s3)) / 250
          + (9 * s4 * (c1 * c2 * s3 + c1 * c3 * s2)) / 250
          - (11 * c1 * c2 * c3) / 50
          + (11 * c1 * s2 * s3) / 50
          - 3.0 / 50.0)
    return np.array([px, py, pz])


def getFootPositionWorld(x_fb, q, biped):
    """World foot positions (6,1), MPC.py:406-424.  Body->world uses R.T (quirk)."""
    rot = eul2rotm(np.asarray(x_fb[0:3], dtype=float))
    out = np.zeros((6, 1))
    hip = np.asarray(biped.hip_offset, dtype=float)
    for leg, side in enumerate((1, -1)):
        pf_b = getFootPositionBody(*q[5 * leg:5 * leg + 5], side)
        offs = np.array([hip[0], side * hip[1], hip[2]])
        out[3 * leg:3 * leg + 3, 0] = np.asarray(x_fb[3:6], dtype=float) + rot.T @ (pf_b + offs)
    return out


def swingLegControl(x_fb, t, pf_w, vf_w, mpc, side):
    """Swing-foot PD force (3,1), MPC.py:426-442."""
    h, dt = mpc.h, mpc.dt
    des_x = x_fb[3] + x_fb[9] * 1 / 2 * h / 2 * dt + mpc.kv * (x_fb[3] - mpc.x_cmd[3])
    des_y = x_fb[4] + x_fb[10] * 1 / 2 * h / 2 * dt + mpc.kv * (x_fb[4] - mpc.x_cmd[4]) + 0.04 * side
    half_cycle = dt * h / 2
    tau = np.remainder(t, half_cycle)
    des_z = mpc.swingHeight * np.sin(np.pi * tau / half_cycle)
    des = np.array([[des_x], [des_y], [des_z]])
    return np.asarray(mpc.kp) @ (des - np.reshape(pf_w, (3, 1))) + np.asarray(mpc.kd) @ (0.0 - np.reshape(vf_w, (3, 1)))


def lowLevelControl(x_fb, t, pf_w, q, qd, mpc, biped, contact, u):
    """Joint torques (10,1) from first-stage forces/moments, MPC.py:444-470."""
    x_fb = np.asarray(x_fb, dtype=float)
    pf_w = np.asarray(pf_w, dtype=float).reshape(6)
    u = np.asarray(u, dtype=float).reshape(12, 1)
    q = np.asarray(q, dtype=float)
    qd = np.asarray(qd, dtype=float)
    c_now = np.asarray(contact, dtype=float)[0, 0:2]
    rot = eul2rotm(x_fb[0:3])
    tau = np.zeros((10, 1))
    for leg, side in enumerate((1, -1)):
        jm, jf = getLegKinematics(*q[5 * leg:5 * leg + 5], side)
        vf_w = rot.T @ jf @ qd[5 * leg:5 * leg + 5].reshape(5, 1)
        f_swing = swingLegControl(x_fb, t, pf_w[3 * leg:3 * leg + 3], vf_w, mpc, side)
        wrench = -np.vstack([rot.T @ u[3 * leg:3 * leg + 3], rot.T @ u[3 * leg + 6:3 * leg + 9]])
        tau[5 * leg:5 * leg + 5] = jm.T @ wrench * c_now[leg] + jf.T @ f_swing * -(c_now[leg] - 1)
    return tau


def mpc_tick(x_fb, t, q, qd, mpc, biped, gait: int = 1, extend: bool = False):
    """One tick in the reference script's call order (MPC.py:475-495).

    Returns dict(states, controls, tau, pf_w, contact, info).
    """
    pf_w = getFootPositionWorld(x_fb, q, biped)
    foot = pf_w.reshape(-1)
    contact = get_contact_sequence(t, mpc, extend=extend) if gait == 1 else np.ones((mpc.h, 2))
    states, controls, info = solve_mpc(x_fb, t, foot, mpc, biped, contact, extend=extend, return_info=True)
    tau = lowLevelControl(x_fb, t, pf_w, q, qd, mpc, biped, contact, controls[0].reshape(-1, 1))
    return dict(states=states, controls=controls, tau=tau, pf_w=pf_w, contact=contact, info=info)


def active_friction_rows(u_stage, contact_stage, mu, scale):
    """Bitmask (8 bits) of non-trivially active friction rows of one stage.

    Row order follows MPC.py:220-229 (+fx,+fy,-fx,-fy per foot).  Per SURVEY.md 7.7 a
    row counts only for a stance foot with fz > 1e-6*scale and residual within
    1e-6*scale of zero, where scale = max(1, |u|_inf) over the whole horizon.
    """
    mask = 0
    tol = 1e-6 * scale
    for leg in range(2):
        if not contact_stage[leg]:
            continue
        fx, fy, fz = u_stage[3 * leg:3 * leg + 3]
        if fz <= tol:
            continue
        res = [fx - mu * fz, fy - mu * fz, -fx - mu * fz, -fy - mu * fz]
        for r, val in enumerate(res):
            if val >= -tol:
                mask |= 1 << (4 * leg + r)
    return mask
