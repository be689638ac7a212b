"""Host-side mirror of the reference's MPC interface over the C ABI.

* :class:`BatchedMPC` - N independent robots per call; device tensors in / device tensors out
  (``step`` / ``solve`` / ``lowlevel`` / ``foot_positions``), or host numpy in / out through one
  packed pinned staging buffer (``step_host`` / ``solve_host``) - the end-to-end path.
* ``solve_mpc`` / ``lowLevelControl`` / ``getFootPositionWorld`` - the reference's own signatures
  (MPC.py:187, 444, 406) for one robot, so the reference script's loop can swap them in.

PyTorch is used only to own device / pinned memory and CUDA streams; every number is produced
by the CUDA kernels in ``csrc/`` behind ``include/biped_mpc_b200.h``.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np

from . import _lib
from .gait import gait_phase
from .params import MPC, Biped, pack_params, params_key

STATUS_OPTIMAL, STATUS_MAXITER, STATUS_NUMERIC, STATUS_BADINPUT = 0, 1, 2, 3


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("biped_mpc_py_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


class BatchedMPC:
    """One solver handle: fixed parameters (``mpc``, ``biped``), one device, batches up to ``max_batch``."""

    def __init__(self, mpc=None, biped=None, max_batch: int = 4096, device: int = 0, extend_gait: bool = False,
                 max_iter: int = 0, mu_tol: float = 0.0, rd_tol: float = 0.0):
        torch = _torch()
        self.mpc = mpc if mpc is not None else MPC()
        self.biped = biped if biped is not None else Biped()
        self.h = int(self.mpc.h)
        self.max_batch = int(max_batch)
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self.extend_gait = bool(extend_gait)
        self._lib = _lib.load()
        self._params = pack_params(self.mpc, self.biped, extend_gait, max_iter, mu_tol, rd_tol)
        handle = ctypes.c_void_p()
        _lib.check(self._lib.bmpc_create(ctypes.byref(self._params), self.device_index, self.max_batch,
                                         ctypes.byref(handle)))
        self._h = handle
        h, nb = self.h, self.max_batch
        f64, dev = torch.float64, self.device
        # outputs are owned by the handle and reused every call (views are returned)
        self.controls = torch.empty((nb, h, 12), dtype=f64, device=dev)
        self.states = torch.empty((nb, h, 13), dtype=f64, device=dev)
        self.tau = torch.empty((nb, 10), dtype=f64, device=dev)
        self.status = torch.empty((nb,), dtype=torch.int32, device=dev)
        self.iters = torch.empty((nb,), dtype=torch.int32, device=dev)
        self.fric_active = torch.empty((nb, h), dtype=torch.uint8, device=dev)
        self.resid = torch.empty((nb, 2), dtype=f64, device=dev)
        self._stage = None  # packed host/device staging, created on first *_host call
        self._ctor = dict(max_iter=max_iter, mu_tol=mu_tol, rd_tol=rd_tol)
        self._options: Dict[str, int] = {}
        self._children: list = []       # per-chunk handles of the chunked host path (ChunkedTick)
        self._warm = None
        self.host_chunk = 65536         # robots per chunk of the chunked host path (step_host for batches of two chunks or more)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        for c in getattr(self, "_children", []):
            c.close()
        self._children = []
        if getattr(self, "_h", None):
            self._lib.bmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.bmpc_launch_count(self._h))

    def set_option(self, name: str, value: int):
        """Dispatch tunables of include/biped_mpc_b200.h::bmpc_set_option (``lane_mode``, ``lane_min``,
        ``lane_ctas_per_sm``, ``lane_warps``, ``lowlat``); every setting returns the same certified optimum."""
        _lib.check(self._lib.bmpc_set_option(self._h, name.encode(), int(value)))
        self._options[name] = int(value)
        for c in self._children:
            c.set_option(name, value)

    def pin_kernel_family(self, family: str = "auto"):
        """Which kernel family solves a robot.  ``"auto"`` (default): decided per call from the batch and class sizes (size
        gates of bmpc.cu::setup_lanes; small batches and small classes run on the warp-per-robot kernels, large ones on the
        lane-per-robot kernels with their later passes) - fastest, but the family, and with it the last bits of the result
        (<= 1e-11 relative, 3e-7 on one robot in a million), depend on how the caller shards its robots.  ``"lane"`` /
        ``"warp"``: every robot takes the same code path whatever the batch it arrives in, so results are bit-identical
        for any sharding or GPU count (``"lane"`` needs the reference's limit structure, five free inputs per foot)."""
        if family not in ("auto", "lane", "warp"):
            raise ValueError("family must be 'auto', 'lane' or 'warp'")
        # -1 = the built-in size gates
        self.set_option("lane_mode", 0 if family == "warp" else -1)
        self.set_option("lane_min", 1 if family == "lane" else -1)
        self.set_option("lane_defer_min", 1 if family == "lane" else -1)
        self.set_option("lowlat", 0 if family != "auto" else -1)

    def warm_start(self, on: bool = True):
        """Warm start across consecutive ``step`` / ``solve`` calls of a caller-owned control loop: robot i of the
        batch must be the same robot one tick later.  Same certified optimum as the cold solve, ~3x faster ticks."""
        _lib.check(self._lib.bmpc_warm_start(self._h, 1 if on else 0))
        self._warm = bool(on)  # (the chunked host path is not used with warm start: the active sets live in this handle)

    def reset_warm_start(self):
        """Forget the stored active sets (after a state reset or a jump in time); warm start stays enabled."""
        _lib.check(self._lib.bmpc_warm_start(self._h, 2))

    def enable_timing(self, on: bool = True):
        _lib.check(self._lib.bmpc_enable_timing(self._h, int(on)))

    def last_timing_ms(self):
        """(classify, lane kernel walking class, lane kernel standing class, warp-per-robot walking, warp-per-robot standing)
        device times of the last tick in ms; while timing is enabled the kernels of a tick run serially."""
        ms = (ctypes.c_float * 5)()
        _lib.check(self._lib.bmpc_last_timing(self._h, ms))
        return [float(v) for v in ms]

    # ------------------------------------------------------------------ helpers
    def _check(self, t, shape, dtype, name):
        torch = _torch()
        if not isinstance(t, torch.Tensor) or t.device != self.device:
            raise TypeError(f"{name}: expected a tensor on {self.device}")
        if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous {dtype} of shape {tuple(shape)}, got {t.dtype} {tuple(t.shape)}")
        return ctypes.c_void_p(t.data_ptr())

    def _stream_ptr(self, stream):
        torch = _torch()
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return ctypes.c_void_p(s.cuda_stream)

    # ------------------------------------------------------------------ device API
    def step(self, x_fb, phase_k, t_swing, foot, contact, q, qd, pf_w, want_states: bool = False, stream=None):
        """solve_mpc + lowLevelControl for N robots (device tensors).  Returns a dict of views."""
        torch = _torch()
        n = int(x_fb.shape[0])
        if n > self.max_batch:
            raise ValueError("batch larger than max_batch")
        f64, h = torch.float64, self.h
        args = [self._check(x_fb, (n, 12), f64, "x_fb"), self._check(phase_k, (n,), torch.int32, "phase_k"),
                self._check(t_swing, (n,), f64, "t_swing"), self._check(foot, (n, 6), f64, "foot"),
                self._check(contact, (n, h, 2), torch.uint8, "contact"), self._check(q, (n, 10), f64, "q"),
                self._check(qd, (n, 10), f64, "qd"), self._check(pf_w, (n, 6), f64, "pf_w")]
        outs = [ctypes.c_void_p(self.controls.data_ptr()),
                ctypes.c_void_p(self.states.data_ptr()) if want_states else ctypes.c_void_p(0),
                ctypes.c_void_p(self.tau.data_ptr()), ctypes.c_void_p(self.status.data_ptr()),
                ctypes.c_void_p(self.iters.data_ptr()), ctypes.c_void_p(self.fric_active.data_ptr()),
                ctypes.c_void_p(self.resid.data_ptr())]
        _lib.check(self._lib.bmpc_step(self._h, n, *args, *outs, self._stream_ptr(stream)))
        return self._views(n, want_states, True)

    def solve(self, x_fb, phase_k, foot, contact, want_states: bool = True, stream=None):
        """solve_mpc only (MPC.py:187-304) for N robots (device tensors)."""
        torch = _torch()
        n = int(x_fb.shape[0])
        f64, h = torch.float64, self.h
        args = [self._check(x_fb, (n, 12), f64, "x_fb"), self._check(phase_k, (n,), torch.int32, "phase_k"),
                self._check(foot, (n, 6), f64, "foot"), self._check(contact, (n, h, 2), torch.uint8, "contact")]
        outs = [ctypes.c_void_p(self.controls.data_ptr()),
                ctypes.c_void_p(self.states.data_ptr()) if want_states else ctypes.c_void_p(0),
                ctypes.c_void_p(self.status.data_ptr()), ctypes.c_void_p(self.iters.data_ptr()),
                ctypes.c_void_p(self.fric_active.data_ptr()), ctypes.c_void_p(self.resid.data_ptr())]
        _lib.check(self._lib.bmpc_solve(self._h, n, *args, *outs, self._stream_ptr(stream)))
        return self._views(n, want_states, False)

    def lowlevel(self, x_fb, t_swing, pf_w, q, qd, contact0, u0, stream=None):
        """lowLevelControl only (MPC.py:444-470): u0 (N,12) -> tau (N,10)."""
        torch = _torch()
        n = int(x_fb.shape[0])
        f64 = torch.float64
        args = [self._check(x_fb, (n, 12), f64, "x_fb"), self._check(t_swing, (n,), f64, "t_swing"),
                self._check(pf_w, (n, 6), f64, "pf_w"), self._check(q, (n, 10), f64, "q"),
                self._check(qd, (n, 10), f64, "qd"), self._check(contact0, (n, 2), torch.uint8, "contact0"),
                self._check(u0, (n, 12), f64, "u0")]
        _lib.check(self._lib.bmpc_lowlevel(self._h, n, *args, ctypes.c_void_p(self.tau.data_ptr()),
                                           self._stream_ptr(stream)))
        return self.tau[:n]

    def foot_positions(self, x_fb, q, out=None, stream=None):
        """getFootPositionWorld (MPC.py:406-424) for N robots -> (N,6)."""
        torch = _torch()
        n = int(x_fb.shape[0])
        f64 = torch.float64
        if out is None:
            out = torch.empty((n, 6), dtype=f64, device=self.device)
        _lib.check(self._lib.bmpc_foot_positions(self._h, n, self._check(x_fb, (n, 12), f64, "x_fb"),
                                                 self._check(q, (n, 10), f64, "q"),
                                                 self._check(out, (n, 6), f64, "out"), self._stream_ptr(stream)))
        return out

    def rollout(self, x, foot, tick, gait, q, qd, ticks: int, warm_start: bool = True, n_log: int = 0, stream=None):
        """Closed loop for N robots, ``ticks`` control ticks (``bmpc_rollout``; rules R1-R6 in DESIGN.md 9).

        ``x`` (N,12), ``foot`` (N,6) and ``tick`` (N,) int32 are device tensors advanced IN PLACE;
        ``gait`` (N,) uint8, ``q``/``qd`` (N,10) are held fixed.  Returns dict(stats=..., and with
        ``n_log`` > 0 the device logs x_log (ticks+1,n_log,12), foot_log, u0_log (ticks,n_log,12), tau_log).
        ``stats`` is a device int64[8] tensor (see include/biped_mpc_b200.h); read it after a sync.
        """
        torch = _torch()
        n = int(x.shape[0])
        if n > self.max_batch:
            raise ValueError("batch larger than max_batch")
        f64, dev = torch.float64, self.device
        args = [self._check(x, (n, 12), f64, "x"), self._check(foot, (n, 6), f64, "foot"),
                self._check(tick, (n,), torch.int32, "tick"), self._check(gait, (n,), torch.uint8, "gait"),
                self._check(q, (n, 10), f64, "q"), self._check(qd, (n, 10), f64, "qd")]
        out = dict(stats=torch.zeros(8, dtype=torch.int64, device=dev))
        null = ctypes.c_void_p(0)
        logs = [null] * 4
        if n_log > 0:
            out["x_log"] = torch.empty((ticks + 1, n_log, 12), dtype=f64, device=dev)
            out["foot_log"] = torch.empty((ticks + 1, n_log, 6), dtype=f64, device=dev)
            out["u0_log"] = torch.empty((ticks, n_log, 12), dtype=f64, device=dev)
            out["tau_log"] = torch.empty((ticks, n_log, 10), dtype=f64, device=dev)
            logs = [ctypes.c_void_p(out[k].data_ptr()) for k in ("x_log", "foot_log", "u0_log", "tau_log")]
        _lib.check(self._lib.bmpc_rollout(self._h, n, int(ticks), *args, int(bool(warm_start)), int(n_log), *logs,
                                          ctypes.c_void_p(out["stats"].data_ptr()), self._stream_ptr(stream)))
        return out

    @staticmethod
    def rollout_stats(stats) -> Dict[str, float]:
        """Name the entries of a rollout ``stats`` tensor (synchronises)."""
        v = [int(a) for a in stats.cpu().tolist()]
        rt = max(1, v[4])
        return dict(robot_ticks=v[4], mean_iters=v[0] / rt, max_iters=v[3], not_optimal=v[1], bad_input=v[2],
                    warm_hits=v[5], warm_hit_rate=v[5] / rt, falls=v[6])

    def debug_assemble(self, x_fb, phase_k, foot, contact):
        """Reduced condensed QP (Hc, g) of ONE instance as numpy arrays (parity tests)."""
        torch = _torch()
        f64, h = torch.float64, self.h
        nmax = 12 * h
        H = torch.zeros((nmax, nmax), dtype=f64, device=self.device)
        g = torch.zeros((nmax,), dtype=f64, device=self.device)
        nn = torch.zeros((1,), dtype=torch.int32, device=self.device)
        _lib.check(self._lib.bmpc_debug_assemble(
            self._h, self._check(x_fb, (1, 12), f64, "x_fb"), self._check(phase_k, (1,), torch.int32, "phase_k"),
            self._check(foot, (1, 6), f64, "foot"), self._check(contact, (1, h, 2), torch.uint8, "contact"),
            ctypes.c_void_p(H.data_ptr()), ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(nn.data_ptr()),
            self._stream_ptr(None)))
        n = int(nn.item())
        return H[:n, :n].cpu().numpy(), g[:n].cpu().numpy()

    def _views(self, n, want_states, with_tau) -> Dict[str, object]:
        out = dict(controls=self.controls[:n], status=self.status[:n], iters=self.iters[:n],
                   fric_active=self.fric_active[:n], resid=self.resid[:n])
        if want_states:
            out["states"] = self.states[:n]
        if with_tau:
            out["tau"] = self.tau[:n]
        return out

    # ------------------------------------------------------------------ host (end-to-end) API
    def _staging(self):
        """One packed pinned host buffer + device mirror for inputs, one for outputs."""
        if self._stage is not None:
            return self._stage
        torch = _torch()
        nb, h = self.max_batch, self.h
        in_bytes = _align(nb * 8 * (12 + 6 + 10 + 10 + 6 + 1)) + _align(nb * 4) + _align(nb * 2 * h) + 7 * 16
        out_bytes = _align(nb * 8 * (12 * h + 13 * h + 10 + 2)) + 2 * _align(nb * 4) + _align(nb * h) + 7 * 16
        st = dict(
            h_in=torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True),
            d_in=torch.empty(in_bytes, dtype=torch.uint8, device=self.device),
            h_out=torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True),
            d_out=torch.empty(out_bytes, dtype=torch.uint8, device=self.device),
        )
        st["h_in_np"] = st["h_in"].numpy()
        st["h_out_np"] = st["h_out"].numpy()
        self._stage = st
        return st

    @staticmethod
    def _carve(sections):
        """sections: list of (name, nbytes) -> dict name -> (offset, nbytes), 16-byte aligned."""
        off, lay = 0, {}
        for name, nbytes in sections:
            lay[name] = (off, nbytes)
            off = _align(off + nbytes)
        return lay, off

    def pinned_tick(self, n: int, lowlevel: bool = True, want_states: bool = False) -> "PinnedTick":
        """Zero-copy host interface: numpy views into pinned staging to fill, then ``run()``."""
        return PinnedTick(self, n, lowlevel, want_states)

    def chunked_tick(self, n: int, chunks: int = 0, lowlevel: bool = True, want_states: bool = False, slot: int = 0) -> "ChunkedTick":
        """Host interface for throughput batches: the batch is cut into ``chunks`` pieces (default: ``host_chunk`` robots each),
        each with its own handle, CUDA stream and packed pinned staging, so that the host->device copy of chunk k+1 and the
        device->host copy of chunk k-1 run under the kernels of chunk k.  Ticks with different ``slot`` numbers own disjoint
        handles and staging: ``launch()`` slot 1 before ``wait()``-ing for slot 0 to pipeline consecutive steps (the copies of
        one step then run under the kernels of the other)."""
        if chunks <= 0:
            chunks = max(1, n // self.host_chunk)
        return ChunkedTick(self, n, chunks, lowlevel, want_states, slot)

    def _chunk_solvers(self, count: int, chunk_n: int, first: int = 0):
        """Child handles ``first .. first+count-1`` for batches up to ``chunk_n`` (same parameters, device and options), each
        with its own stream."""
        torch = _torch()
        if any(c.max_batch < chunk_n for c in self._children[first:first + count]):
            self._children = self._drop_children()
        while len(self._children) < first + count:
            c = BatchedMPC(self.mpc, self.biped, max_batch=chunk_n, device=self.device_index, extend_gait=self.extend_gait,
                           **self._ctor)
            c._own_stream = torch.cuda.Stream(device=self.device)
            for k, v in self._options.items():
                c.set_option(k, v)
            self._children.append(c)
        return self._children[first:first + count]

    def _drop_children(self):
        for c in self._children:
            c.close()
        return []

    def step_host(self, x_fb, t, foot, contact, q, qd, pf_w, phase_k=None, want_states: bool = False,
                  lowlevel: bool = True):
        """End-to-end tick with HOST numpy inputs and outputs (one H2D and one D2H copy).

        ``t`` (N,) is the control time: ``phase_k`` defaults to ``int(t // dt) % h`` computed with the
        reference's float expression (MPC.py:56,99) and ``t`` is also the swing-leg time (MPC.py:436).
        Returns dict of numpy arrays: controls (N,h,12), tau (N,10), status, iters, fric_active, resid
        and optionally states (N,h,13).  Synchronous.
        """
        x_fb = np.ascontiguousarray(x_fb, dtype=np.float64).reshape(-1, 12)
        n, h = x_fb.shape[0], self.h
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(n)
        if phase_k is None:
            period = 10 if self.extend_gait else h
            phase_k = (gait_phase(t, self.mpc) % period)
        arrays = dict(x_fb=x_fb, foot=np.asarray(foot, dtype=np.float64).reshape(n, 6),
                      phase_k=np.asarray(phase_k, dtype=np.int32).reshape(n),
                      contact=np.asarray(contact, dtype=np.uint8).reshape(n, h, 2))
        if lowlevel:
            arrays.update(q=np.asarray(q, dtype=np.float64).reshape(n, 10), qd=np.asarray(qd, dtype=np.float64).reshape(n, 10),
                          pf_w=np.asarray(pf_w, dtype=np.float64).reshape(n, 6), t=t)
        if n >= 2 * self.host_chunk and not self._warm:
            # throughput batch: chunked, copies overlapped with the kernels of the neighbouring chunks
            tick = self.chunked_tick(n, 0, lowlevel, want_states)
            tick.set_inputs(**arrays)
            tick.run()
            self.last_h2d_bytes, self.last_d2h_bytes = tick.h2d_bytes, tick.d2h_bytes
            return tick.gather()
        tick = self.pinned_tick(n, lowlevel, want_states)
        for k, a in arrays.items():
            tick.inputs[k][...] = a
        out = tick.run()
        self.last_h2d_bytes, self.last_d2h_bytes = tick.h2d_bytes, tick.d2h_bytes
        return {k: v.copy() for k, v in out.items()}

    def tick_host(self, x_fb, t, q, qd, gait, want_states: bool = False):
        """One control tick for N robots in the reference script's call order (MPC.py:475-495): forward kinematics
        ``pf_w = getFootPositionWorld(x_fb, q)`` and ``foot = pf_w`` (MPC.py:478-479), the contact schedule from ``t`` and
        ``gait`` (1 walking: ``get_contact_sequence``, 0 standing: ones; MPC.py:481-484), ``solve_mpc`` and
        ``lowLevelControl`` on ``controls[0]`` (MPC.py:487-494).  Host numpy in / out; FK, solve and torque map run on the
        device, the gait table and the float phase (MPC.py:56) on the host.  Returns the ``step_host`` dict plus ``pf_w``
        and ``contact``."""
        torch = _torch()
        from .gait import batch_contact_and_phase
        x_fb = np.ascontiguousarray(x_fb, dtype=np.float64).reshape(-1, 12)
        n = x_fb.shape[0]
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(n, 10)
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(n)
        gait = np.broadcast_to(np.asarray(gait), (n,))
        dev = self.device
        pf_w = self.foot_positions(torch.as_tensor(x_fb, device=dev), torch.as_tensor(q, device=dev)).cpu().numpy()
        contact, phase_k = batch_contact_and_phase(t, gait, self.mpc, extend=self.extend_gait)
        out = self.step_host(x_fb, t, pf_w, contact, q, qd, pf_w, phase_k=phase_k, want_states=want_states)
        out["pf_w"], out["contact"] = pf_w, contact
        return out

    def solve_host(self, x_fb, t, foot, contact, phase_k=None, want_states: bool = True):
        """``solve_mpc`` only, host numpy in / out."""
        return self.step_host(x_fb, t, foot, contact, None, None, None, phase_k=phase_k, want_states=want_states,
                              lowlevel=False)


class PinnedTick:
    """One batch size's packed pinned layout: fill ``inputs`` (numpy views), call ``run()``.

    ``run()`` = one ``cudaMemcpyAsync`` host->device of the packed inputs, the fused kernels, one
    device->host copy of the packed outputs, a stream synchronise; it returns numpy views into the
    pinned output buffer (valid until the next ``run()`` of any tick of the same solver).
    """

    def __init__(self, solver: BatchedMPC, n: int, lowlevel: bool, want_states: bool):
        if n > solver.max_batch or n <= 0:
            raise ValueError("batch size out of range")
        self.s, self.n, self.lowlevel, self.want_states = solver, n, lowlevel, want_states
        h = solver.h
        st = solver._staging()
        ins = [("x_fb", np.float64, (n, 12)), ("foot", np.float64, (n, 6))]
        if lowlevel:
            ins += [("q", np.float64, (n, 10)), ("qd", np.float64, (n, 10)), ("pf_w", np.float64, (n, 6)),
                    ("t", np.float64, (n,))]
        ins += [("phase_k", np.int32, (n,)), ("contact", np.uint8, (n, h, 2))]
        outs = [("controls", np.float64, (n, h, 12))]
        if lowlevel:
            outs.append(("tau", np.float64, (n, 10)))
        if want_states:
            outs.append(("states", np.float64, (n, h, 13)))
        outs += [("resid", np.float64, (n, 2)), ("status", np.int32, (n,)), ("iters", np.int32, (n,)),
                 ("fric_active", np.uint8, (n, h))]
        nbytes = lambda dt, shp: int(np.dtype(dt).itemsize * int(np.prod(shp)))
        self.lay_in, self.h2d_bytes = BatchedMPC._carve([(k, nbytes(dt, shp)) for k, dt, shp in ins])
        self.lay_out, self.d2h_bytes = BatchedMPC._carve([(k, nbytes(dt, shp)) for k, dt, shp in outs])
        view = lambda buf, lay, k, dt, shp: buf[lay[k][0]:lay[k][0] + lay[k][1]].view(dt).reshape(shp)
        self.inputs = {k: view(st["h_in_np"], self.lay_in, k, dt, shp) for k, dt, shp in ins}
        self.outputs = {k: view(st["h_out_np"], self.lay_out, k, dt, shp) for k, dt, shp in outs}
        self._st = st

    def run(self):
        stream = self.launch()
        stream.synchronize()
        return self.outputs

    def launch(self):
        """Enqueue host->device copy, kernels and device->host copy on the current stream; returns the stream."""
        torch = _torch()
        s, st, n = self.s, self._st, self.n
        stream = torch.cuda.current_stream(s.device)
        st["d_in"][:self.h2d_bytes].copy_(st["h_in"][:self.h2d_bytes], non_blocking=True)
        din, dout = st["d_in"].data_ptr(), st["d_out"].data_ptr()
        I = lambda k: ctypes.c_void_p(din + self.lay_in[k][0])
        O = lambda k: ctypes.c_void_p(dout + self.lay_out[k][0]) if k in self.lay_out else ctypes.c_void_p(0)
        sp = ctypes.c_void_p(stream.cuda_stream)
        if self.lowlevel:
            _lib.check(s._lib.bmpc_step(s._h, n, I("x_fb"), I("phase_k"), I("t"), I("foot"), I("contact"), I("q"),
                                        I("qd"), I("pf_w"), O("controls"), O("states"), O("tau"), O("status"),
                                        O("iters"), O("fric_active"), O("resid"), sp))
        else:
            _lib.check(s._lib.bmpc_solve(s._h, n, I("x_fb"), I("phase_k"), I("foot"), I("contact"), O("controls"),
                                         O("states"), O("status"), O("iters"), O("fric_active"), O("resid"), sp))
        st["h_out"][:self.d2h_bytes].copy_(st["d_out"][:self.d2h_bytes], non_blocking=True)
        return stream


class ChunkedTick:
    """A throughput batch from HOST memory in ``chunks`` pieces.  Every chunk has its own solver handle (hence its own work
    lists and workspace), its own CUDA stream and its own packed pinned staging, and is enqueued as one host->device
    ``cudaMemcpyAsync`` of its packed inputs, the fused kernels, and one device->host ``cudaMemcpyAsync`` of its packed
    outputs.  The chunks' streams are independent, so the copies of chunk k+1 / k-1 run under the kernels of chunk k and
    the kernels of consecutive chunks fill the SMs back to back (the persistent CTAs of chunk k+1 start as those of chunk k
    run out of work).  Robots are independent (MPC.py:187-304 has no shared state), so chunking changes no result."""

    def __init__(self, solver: BatchedMPC, n: int, chunks: int, lowlevel: bool, want_states: bool, slot: int = 0):
        if n <= 0 or n > solver.max_batch or chunks < 1 or slot < 0:
            raise ValueError("batch size / chunk count / slot out of range")
        from .shard import shard_slice
        self.s, self.n = solver, n
        self.slices = [shard_slice(n, i, chunks) for i in range(chunks)]
        chunk_n = max(sl.stop - sl.start for sl in self.slices)
        self.children = solver._chunk_solvers(chunks, chunk_n, first=slot * chunks)
        self.parts = [PinnedTick(c, sl.stop - sl.start, lowlevel, want_states) for c, sl in zip(self.children, self.slices)]
        self.h2d_bytes = sum(p.h2d_bytes for p in self.parts)
        self.d2h_bytes = sum(p.d2h_bytes for p in self.parts)

    def set_inputs(self, **arrays):
        """Copy whole-batch host arrays (x_fb, foot, phase_k, contact, and with lowlevel q, qd, pf_w, t) into the chunks' pinned staging."""
        for part, sl in zip(self.parts, self.slices):
            for k, a in arrays.items():
                part.inputs[k][...] = a[sl]

    def launch(self):
        torch = _torch()
        for part, c in zip(self.parts, self.children):
            with torch.cuda.stream(c._own_stream):
                part.launch()

    def wait(self):
        """Block until every chunk of the last ``launch()`` has finished; returns the per-chunk output views (pinned memory,
        valid until the next launch of this tick)."""
        for c in self.children:
            c._own_stream.synchronize()
        return [p.outputs for p in self.parts]

    def run(self):
        """``launch()`` + ``wait()``."""
        self.launch()
        return self.wait()

    def gather(self):
        """Whole-batch numpy copies of the outputs of the last ``run()``."""
        return {k: np.concatenate([p.outputs[k] for p in self.parts], axis=0) for k in self.parts[0].outputs}


# ----------------------------------------------------------------------------------------------
# the reference's single-robot signatures
# ----------------------------------------------------------------------------------------------
_SOLVERS: Dict[bytes, BatchedMPC] = {}


def default_solver(mpc, biped, max_batch: int = 1, extend_gait: bool = False) -> BatchedMPC:
    """Cached handle for a parameter set (the shim functions are stateless like the reference's)."""
    key = params_key(pack_params(mpc, biped, extend_gait)) + bytes([max_batch > 1])
    s = _SOLVERS.get(key)
    if s is None:
        s = BatchedMPC(mpc, biped, max_batch=max(1, max_batch), extend_gait=extend_gait)
        _SOLVERS[key] = s
    return s


def _check_status(who: str, status: int):
    """The reference never looks at its solver's status (MPC.py:297-300); here a result that is not the certified optimum is
    never passed on silently: bad input raises, an uncertified iterate warns."""
    if status == STATUS_BADINPUT:
        raise ValueError(f"{who}: non-finite input or singular euler-rate matrix (pitch = +-pi/2)")
    if status != STATUS_OPTIMAL:
        import warnings
        warnings.warn(f"{who}: the returned point is an interior-point iterate that was NOT certified optimal (status {status})",
                      RuntimeWarning, stacklevel=3)


def solve_mpc(x_fb, t, foot, mpc, biped, contact):
    """Drop-in for ``solve_mpc`` (MPC.py:187-304): returns ``(states (h,13), controls (h,12))``.

    Differences from the reference, on purpose: nothing is printed (MPC.py:190-192 prints the
    references each call); the result is the optimum of the QP the reference builds, to ~1e-7
    relative, rather than cvxopt's iterate at its default tolerances.
    """
    h = int(mpc.h)
    contact = np.asarray(contact)
    if contact.shape != (h, 2):
        raise IndexError(f"contact must have shape ({h}, 2), got {contact.shape} "
                         "(the reference fails the same way for h != 10 walking, MPC.py:58)")
    s = default_solver(mpc, biped)
    args = (np.asarray(x_fb, dtype=np.float64).reshape(1, 12), np.array([float(t)]),
            np.asarray(foot, dtype=np.float64).reshape(1, 6), (contact != 0).astype(np.uint8)[None])
    out = s.solve_host(*args)
    if int(out["status"][0]) in (STATUS_MAXITER, STATUS_NUMERIC):
        # the latency path (batches of 1 .. 8) skips the last-resort pass of the library: repeat the instance in a batch that has it
        out = default_solver(mpc, biped, max_batch=16).solve_host(*[np.repeat(a, 9, axis=0) for a in args])
    _check_status("solve_mpc", int(out["status"][0]))
    return out["states"][0], out["controls"][0]


def lowLevelControl(x_fb, t, pf_w, q, qd, mpc, biped, contact, u):
    """Drop-in for ``lowLevelControl`` (MPC.py:444-470): returns tau (10,1)."""
    torch = _torch()
    s = default_solver(mpc, biped)
    dev, f64 = s.device, torch.float64
    c0 = (np.asarray(contact)[0, 0:2] != 0).astype(np.uint8).reshape(1, 2)
    tn = lambda a, shape: torch.as_tensor(np.asarray(a, dtype=np.float64).reshape(shape), dtype=f64, device=dev)
    tau = s.lowlevel(tn(x_fb, (1, 12)), tn([float(t)], (1,)), tn(pf_w, (1, 6)), tn(q, (1, 10)), tn(qd, (1, 10)),
                     torch.as_tensor(c0, device=dev), tn(u, (1, 12)))
    return tau[0].cpu().numpy().reshape(10, 1)


def getFootPositionWorld(x_fb, q, biped, mpc=None):
    """Drop-in for ``getFootPositionWorld`` (MPC.py:406-424): returns pf_w (6,1)."""
    torch = _torch()
    s = default_solver(mpc if mpc is not None else MPC(), biped)
    tn = lambda a, shape: torch.as_tensor(np.asarray(a, dtype=np.float64).reshape(shape), dtype=torch.float64,
                                          device=s.device)
    return s.foot_positions(tn(x_fb, (1, 12)), tn(q, (1, 10)))[0].cpu().numpy().reshape(6, 1)


def mpc_tick(x_fb, t, q, qd, mpc, biped, gait: int = 1):
    """The reference's main script as a function (MPC.py:475-495), one robot: FK, contact schedule, ``solve_mpc``,
    ``lowLevelControl``.  Returns dict(states (h,13), controls (h,12), tau (10,1), pf_w (6,1), contact (h,2)).
    For a control loop call ``default_solver(mpc, biped).warm_start(True)`` once before the first tick."""
    s = default_solver(mpc, biped)
    out = s.tick_host(np.asarray(x_fb, dtype=np.float64).reshape(1, 12), np.array([float(t)]), np.asarray(q).reshape(1, 10),
                      np.asarray(qd).reshape(1, 10), np.array([int(gait)]), want_states=True)
    if int(out["status"][0]) in (STATUS_MAXITER, STATUS_NUMERIC):  # (see solve_mpc)
        rep9 = lambda a: np.repeat(np.asarray(a).reshape(1, -1), 9, axis=0)
        out = default_solver(mpc, biped, max_batch=16).tick_host(rep9(x_fb), np.full(9, float(t)), rep9(q), rep9(qd), np.full(9, int(gait)),
                                                                 want_states=True)
    _check_status("mpc_tick", int(out["status"][0]))
    return dict(states=out["states"][0], controls=out["controls"][0], tau=out["tau"][0].reshape(10, 1),
                pf_w=out["pf_w"][0].reshape(6, 1), contact=out["contact"][0])
