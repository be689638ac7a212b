"""biped_mpc_py_b200 - B200-native batched solver for the HECTOR-style force-and-moment MPC.

Drop-in for the hot path of junhengl/biped_mpc_py (``bipedalLocomotionMPC.py``):
``solve_mpc`` (MPC.py:187-304) and ``lowLevelControl`` (MPC.py:444-470), one robot through the
reference's own signatures or thousands through :class:`BatchedMPC`.  All numerics run in
hand-written sm_100a CUDA kernels behind a C ABI (``include/biped_mpc_b200.h``); there is no
CPU fallback - importing works anywhere, computing requires a B200 and the built library.
"""
from .params import MPC, Biped, pack_params
from .gait import gait_phase, get_contact_sequence
from .api import BatchedMPC, solve_mpc, lowLevelControl, getFootPositionWorld, default_solver, mpc_tick
from .sim import SimulatorAdapter
from . import synth

__all__ = ["MPC", "Biped", "pack_params", "gait_phase", "get_contact_sequence", "BatchedMPC", "solve_mpc",
           "lowLevelControl", "getFootPositionWorld", "default_solver", "mpc_tick", "SimulatorAdapter", "synth"]
