"""Host-side gait scheduler: produces the ``contact`` and ``phase_k`` inputs of the solver.

The integer phase must be computed with the reference's *floating-point* floor division
(MPC.py:56, 99): ``int((3*0.04)//0.04) == 2``, off by one at many tick boundaries.  It is
therefore computed here, on the host, with the same expression and handed to the kernels as
an integer.
"""
from __future__ import annotations

import numpy as np

GAIT_PERIOD = 10  # 5 ticks left stance, 5 ticks right stance (MPC.py:52-55)


def gait_phase(t, mpc):
    """``int(t // mpc.dt)`` evaluated in float64 like MPC.py:56 (scalar or array ``t``)."""
    ph = np.floor_divide(np.asarray(t, dtype=np.float64), float(mpc.dt))
    return ph.astype(np.int64) if ph.ndim else int(ph)


def _table():
    left = (np.arange(2 * GAIT_PERIOD) % GAIT_PERIOD) < GAIT_PERIOD // 2
    return np.stack([left, ~left], axis=1).astype(np.uint8)


def get_contact_sequence(t, mpc, extend: bool = False):
    """Contact schedule of MPC.py:50-59 for one ``t`` -> (rows, 2) uint8.

    Reference behaviour: rows ``k:k+10`` of the 20-row table, ``k = phase % h`` (always 10 rows;
    for h != 10 the reference then fails).  ``extend=True`` continues the table periodically to
    ``h`` rows (identical for h = 10).
    """
    phase = gait_phase(t, mpc)
    table = _table()
    if not extend:
        k = phase % int(mpc.h)
        return table[k:k + 10, :]
    k = phase % GAIT_PERIOD
    return table[(k + np.arange(int(mpc.h))) % GAIT_PERIOD, :]


def batch_contact_and_phase(t, gait, mpc, extend: bool = False):
    """Vectorised: ``t`` (N,), ``gait`` (N,) 1 = walking / 0 = standing -> contact (N,h,2) uint8, phase_k (N,) int32.

    ``phase_k`` is ``phase % h`` (``phase % 10`` with ``extend``), the ``k`` of MPC.py:100.
    """
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    gait = np.asarray(gait).reshape(-1)
    h = int(mpc.h)
    phase = gait_phase(t, mpc)
    period = GAIT_PERIOD if (extend or h == GAIT_PERIOD) else h
    k = (phase % period).astype(np.int64)
    if h != GAIT_PERIOD and not extend:
        raise ValueError("walking gait with h != 10 needs extend=True (the reference raises IndexError here)")
    rows = (k[:, None] + np.arange(h)[None, :]) % GAIT_PERIOD
    contact = _table()[rows]
    contact = np.where(gait[:, None, None] == 1, contact, np.uint8(1)).astype(np.uint8)
    return np.ascontiguousarray(contact), k.astype(np.int32)
