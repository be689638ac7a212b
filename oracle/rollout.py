"""Closed-loop single-rigid-body rollout on the CPU: the checker of ``bmpc_rollout``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference has no loop and no simulator (MPC.py:475-495 runs ONE tick), so the closed loop is
new behaviour defined from the reference's own pieces; the product's
``bmpc_rollout`` follows exactly these rules and ``DESIGN.md`` states them:

R1  clock      The tick counter T is an integer; the gait phase is T itself (``k = T % 10``,
               MPC.py:56-58 without the float floor-division, which only applies when time
               arrives as a float) and the swing-leg clock is ``t = T*dt`` (MPC.py:436).
R2  contact    walking: rows ``(k+i) % 10`` of the table at MPC.py:52-55; standing: ones
               (MPC.py:482-484).
R3  feet       ``foot`` (MPC.py:479) is rollout state and ``pf_w = foot``: stance feet stay
               where they are; a swing foot keeps its last position until touchdown.
R4  tick       ``solve_mpc`` then ``lowLevelControl`` on ``controls[0]`` (MPC.py:487-494);
               q, qd are held at their initial values (they only enter tau).
R5  plant      ``x+ = A_0 [x;1] + B_0 u_0`` with A_0, B_0 the reference's own discretisation at
               the current state and feet (MPC.py:148-185, 206-208): the model is the plant.
R6  touchdown  a foot that is in swing at tick T and in stance at tick T+1 is placed, at the
               NEW state x+, on the swing controller's target (MPC.py:427-435) with z = 0.
R7  fall       if x+ is not finite, any |euler| > 0.8 rad or z outside [0.25, 1.0] m, the robot has
               fallen: it is put back on the reference's initial state (MPC.py:13) with
               ``foot = getFootPositionWorld`` of that state and its own q (MPC.py:406-424); the clock
               keeps running.  (The reference formulation is not globally stabilising even with
               model = plant: about 1 robot in 4,000 tips over within 300 ticks from the perturbed
               nominal state; beyond the thresholds it never recovers and its QP data overflow.)
"""
from __future__ import annotations

import numpy as np

from . import reference_mpc as rm

GAIT_PERIOD = rm.GAIT_PERIOD


def contact_rows(tick: int, gait: int, h: int):
    """R2: (h,2) contact schedule seen at integer tick ``tick``."""
    if gait == 0:
        return np.ones((h, 2), dtype=int)
    left = (np.arange(GAIT_PERIOD) < GAIT_PERIOD // 2)
    table = np.stack([left, ~left], axis=1).astype(int)
    return table[(tick % GAIT_PERIOD + np.arange(h)) % GAIT_PERIOD, :]


def srb_step(x, foot, u0, mpc, biped):
    """R5: one step of the reference's discretised single-rigid-body model (MPC.py:148-185)."""
    xa = np.concatenate([np.asarray(x, dtype=float).reshape(12), [1.0]])
    a_mat, b_mat = rm.get_simplified_dynamics(mpc, biped, xa, np.asarray(foot, dtype=float).reshape(6))
    return (a_mat @ xa + b_mat @ np.asarray(u0, dtype=float).reshape(12))[:12]


def touchdown_target(x, mpc, side):
    """R6: x,y of the swing controller's target (MPC.py:427-435), z = 0."""
    fx = x[3] + x[9] * 1 / 2 * mpc.h / 2 * mpc.dt + mpc.kv * (x[3] - mpc.x_cmd[3])
    fy = x[4] + x[10] * 1 / 2 * mpc.h / 2 * mpc.dt + mpc.kv * (x[4] - mpc.x_cmd[4]) + 0.04 * side
    return np.array([fx, fy, 0.0])


FALL_EULER, FALL_Z_LO, FALL_Z_HI = 0.8, 0.25, 1.0


def fallen(x) -> bool:
    """R7 test."""
    x = np.asarray(x, dtype=float)
    return bool((not np.isfinite(x).all()) or np.abs(x[0:3]).max() > FALL_EULER or x[5] < FALL_Z_LO or x[5] > FALL_Z_HI)


def advance(x, foot, tick, gait, u0, mpc, biped, q=None):
    """R5 + R6 + R7 + clock: returns (x+, foot+, tick+1).  ``q`` is needed only when the robot falls."""
    xn = srb_step(x, foot, u0, mpc, biped)
    if fallen(xn):
        xr = rm.X_FB0.copy()
        return xr, rm.getFootPositionWorld(xr, np.asarray(q, dtype=float), biped).reshape(-1), tick + 1
    fn = np.asarray(foot, dtype=float).reshape(6).copy()
    if gait == 1:
        now = contact_rows(tick, gait, 1)[0]
        nxt = contact_rows(tick + 1, gait, 1)[0]
        for leg, side in enumerate((1.0, -1.0)):
            if nxt[leg] == 1 and now[leg] == 0:
                fn[3 * leg:3 * leg + 3] = touchdown_target(xn, mpc, side)
    return xn, fn, tick + 1


def tick_once(x, foot, tick, gait, q, qd, mpc, biped):
    """R1-R4 for one robot: returns dict(controls, states, tau, contact)."""
    h = mpc.h
    contact = contact_rows(tick, gait, h)
    t_phase = (tick + 0.5) * mpc.dt  # int(t_phase // dt) == tick exactly: the integer clock of R1
    states, controls = rm.solve_mpc(x, t_phase, foot, mpc, biped, contact, extend=True)
    tau = rm.lowLevelControl(x, tick * mpc.dt, np.asarray(foot, dtype=float).reshape(6, 1), q, qd, mpc, biped,
                             contact, controls[0].reshape(-1, 1)).reshape(-1)
    return dict(controls=controls, states=states, tau=tau, contact=contact)


def rollout(x0, foot0, tick0, gait, q, qd, ticks, mpc, biped):
    """Closed loop for ONE robot: returns dict of per-tick logs x (ticks+1,12), foot (ticks+1,6), u0, tau."""
    x = np.asarray(x0, dtype=float).reshape(12).copy()
    foot = np.asarray(foot0, dtype=float).reshape(6).copy()
    tick = int(tick0)
    xs, fs, us, taus = [x.copy()], [foot.copy()], [], []
    for _ in range(ticks):
        out = tick_once(x, foot, tick, gait, q, qd, mpc, biped)
        u0 = out["controls"][0]
        us.append(u0.copy())
        taus.append(out["tau"].copy())
        x, foot, tick = advance(x, foot, tick, gait, u0, mpc, biped, q)
        xs.append(x.copy())
        fs.append(foot.copy())
    return dict(x=np.array(xs), foot=np.array(fs), u0=np.array(us), tau=np.array(taus))
