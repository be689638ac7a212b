"""Step/observe shim between a physics simulator and the batched controller.

The reference announces a simulator interface ("MuJoCo interface is on the way", README.md:7) and its main script
(MPC.py:475-495) shows what one control period of it does for one robot: read the state and the joint angles, compute the
foot positions by forward kinematics (MPC.py:478-479), pick the contact schedule of the gait (MPC.py:481-484), call
``solve_mpc`` (MPC.py:487) and ``lowLevelControl`` on ``controls[0]`` (MPC.py:493-494), hand the joint torques back.
:class:`SimulatorAdapter` is that loop body as an object, for N robots at once: the simulator owns the physics and the
sensors, the adapter owns the controller's clock, gait schedule and warm-start store.

    sim = MySimulator(n)                                  # MuJoCo, Isaac, a plant model ...
    ctl = SimulatorAdapter(n, mpc, biped, gait=1)
    while sim.running():
        ctl.observe(sim.base_state(), sim.q(), sim.qd())  # (n,12), (n,10), (n,10): numpy or tensors on ctl.device
        sim.apply_torques(ctl.step())                     # (n,10), same kind as the observation
        sim.advance(mpc.dt)

There is no simulator in the reference or in this image, so the simulator side is the caller's; everything on the
controller side of the boundary runs on the device (forward kinematics, solve, torque map) except the gait table, which is a
ten-row lookup on the host.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .api import BatchedMPC, STATUS_BADINPUT, _torch
from .gait import batch_contact_and_phase
from .params import MPC, Biped


class SimulatorAdapter:
    """One controller for ``n`` simulated robots.

    ``gait``: 1 walking / 0 standing, a scalar or one value per robot (MPC.py:481-484).  ``warm_start``: start every tick's
    active-set polish from the previous tick's certified active set (same optimum, ~2.5x less latency; the reference solves
    cold every call, MPC.py:297).  ``strict``: raise if a tick returns a robot that is not certified optimal instead of
    only reporting it in ``last["status"]``.
    """

    def __init__(self, n: int, mpc=None, biped=None, gait=1, device: int = 0, warm_start: bool = True, strict: bool = False):
        self.mpc = mpc if mpc is not None else MPC()
        self.biped = biped if biped is not None else Biped()
        self.n = int(n)
        self.solver = BatchedMPC(self.mpc, self.biped, max_batch=self.n, device=device)
        self.device = self.solver.device
        self.strict = bool(strict)
        self.gait = np.ascontiguousarray(np.broadcast_to(np.asarray(gait, dtype=np.int32), (self.n,)))
        self.tick = np.zeros(self.n, dtype=np.int64)   # control periods since the robot's last reset
        self._warm = bool(warm_start)
        if self._warm:
            self.solver.warm_start(True)
        self._obs = None
        self._numpy_io = True
        self.last: Dict[str, object] = {}

    # ---- simulator -> controller ---------------------------------------------------------------------------------
    def observe(self, x_fb, q, qd, t=None):
        """What the simulator measured at the start of this control period: base state ``x_fb`` (n,12) in the reference's
        order [euler, position, angular velocity, velocity] (MPC.py:13), joint angles and rates (n,10) each.  ``t`` (n,):
        the simulator's clock; by default the adapter counts control periods itself (``t = tick * dt``)."""
        torch = _torch()
        self._numpy_io = not torch.is_tensor(x_fb)
        as_dev = lambda a, cols: torch.as_tensor(a, dtype=torch.float64, device=self.device).reshape(self.n, cols).contiguous()
        self._obs = (as_dev(x_fb, 12), as_dev(q, 10), as_dev(qd, 10),
                     None if t is None else np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(self.n)))

    # ---- controller -> simulator ---------------------------------------------------------------------------------
    def step(self, want_states: bool = False):
        """One control period on the last observation: joint torques (n,10) to hold for ``mpc.dt``.  Details of the tick
        (controls, predicted states, contact schedule, foot positions, status, iterations) are in ``self.last``."""
        if self._obs is None:
            raise RuntimeError("SimulatorAdapter.step: call observe() first")
        torch = _torch()
        x_fb, q, qd, t = self._obs
        self._obs = None
        dt = float(self.mpc.dt)
        if t is None:
            # the adapter's own clock is an integer: the gait phase is the tick itself (no float floor division, which is off by
            # one at 27 of the first 60 period boundaries, MPC.py:56) and the swing clock is tick * dt (MPC.py:436)
            t = self.tick.astype(np.float64) * dt
            period = 10
            phase_k = (self.tick % period).astype(np.int32)
            rows = (phase_k[:, None] + np.arange(int(self.mpc.h))[None, :]) % period      # rows of the table at MPC.py:52-55
            left = rows < period // 2
            contact = np.where((self.gait == 1)[:, None, None], np.stack([left, ~left], axis=2), True).astype(np.uint8)
        else:
            contact, phase_k = batch_contact_and_phase(t, self.gait, self.mpc)   # the reference's float expression
        pf_w = self.solver.foot_positions(x_fb, q)                                # MPC.py:478
        dev = self.device
        out = self.solver.step(x_fb, torch.as_tensor(np.ascontiguousarray(phase_k, dtype=np.int32), device=dev),
                               torch.as_tensor(t, dtype=torch.float64, device=dev), pf_w,          # foot = pf_w, MPC.py:479
                               torch.as_tensor(np.ascontiguousarray(contact, dtype=np.uint8), device=dev), q, qd, pf_w,
                               want_states=want_states)
        self.tick += 1
        status = out["status"].cpu().numpy()
        if self.strict and (status != 0).any():
            bad = np.nonzero(status != 0)[0]
            kind = "non-finite input or pitch = +-pi/2" if (status[bad] == STATUS_BADINPUT).any() else "not certified optimal"
            raise RuntimeError(f"SimulatorAdapter.step: robots {bad[:8].tolist()} {kind} (status {status[bad][:8].tolist()})")
        conv = (lambda a: a.cpu().numpy()) if self._numpy_io else (lambda a: a)
        self.last = dict(controls=conv(out["controls"]), tau=conv(out["tau"]), pf_w=conv(pf_w), contact=contact,
                         status=status, iters=out["iters"].cpu().numpy(), t=t)
        if want_states:
            self.last["states"] = conv(out["states"])
        return self.last["tau"]

    def reset(self, robots: Optional[np.ndarray] = None):
        """The simulator put some robots (default: all) back to an initial state: their clocks restart; the warm-start store is
        forgotten (the next tick solves cold, as after construction)."""
        if robots is None:
            self.tick[:] = 0
        else:
            self.tick[np.asarray(robots)] = 0
        if self._warm:
            self.solver.reset_warm_start()

    def close(self):
        self.solver.close()
