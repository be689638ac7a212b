"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares (no compute call is made - there is no GPU here), the ctypes struct matches the header's
field order, and the host-side mirror (gait table, parameter packing) follows the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "biped_mpc_b200.h")


def _header_text():
    with open(HEADER) as f:
        return f.read()


def _declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", _header_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(bmpc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from biped_mpc_py_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from biped_mpc_py_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == declared, "ctypes binding and header disagree on the symbol list"
    assert lib.bmpc_abi_version() == int(re.search(r"#define BMPC_ABI_VERSION (\d+)", _header_text()).group(1))


def test_params_struct_matches_header_order():
    from biped_mpc_py_b200.params import BmpcParams
    body = re.search(r"typedef struct bmpc_params \{(.*?)\} bmpc_params;", _header_text(), flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(",")[0:1] + decl.split(",")[1:]:
            m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[\d+\])?\s*$", part.strip())
            names.append(m.group(1))
    assert names == [f[0] for f in BmpcParams._fields_]


def test_no_gpu_means_loud_failure_not_fallback(lib):
    """Without a CUDA device bmpc_create must fail with a message; the Python API must raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from biped_mpc_py_b200 import MPC, Biped, pack_params, BatchedMPC
    h = ctypes.c_void_p()
    rc = lib.bmpc_create(ctypes.byref(pack_params(MPC(), Biped())), 0, 16, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.bmpc_last_error()
    with pytest.raises(RuntimeError):
        BatchedMPC(MPC(), Biped(), max_batch=4)
    from biped_mpc_py_b200 import SimulatorAdapter
    with pytest.raises(RuntimeError):      # the simulator adapter has no CPU path either
        SimulatorAdapter(2, MPC(), Biped())


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "biped_mpc_py_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_gait_mirror_follows_reference_float_phase():
    from biped_mpc_py_b200 import MPC, get_contact_sequence, gait_phase
    from oracle import reference_mpc as rm
    mpc = MPC()
    for t in np.arange(0, 2.4, 0.04).tolist() + [0.12, 0.28, 0.7999999]:
        assert int(gait_phase(np.array([t]), mpc)[0]) == rm.gait_phase(t, rm.MPCParams())
        assert np.array_equal(get_contact_sequence(t, mpc), rm.get_contact_sequence(t, rm.MPCParams()))
    assert int((3 * 0.04) // 0.04) == 2  # the quirk itself (MPC.py:56)


def test_synthetic_workload_feet_match_the_oracle_kinematics():
    """The synthetic workload's foot positions (biped_mpc_py_b200/synth.py, a vectorised host mirror of getFootPositionWorld,
    MPC.py:406-424) against the oracle, pose by pose: the bench and the property tests start from the reference's kinematics."""
    from biped_mpc_py_b200 import synth
    from oracle import reference_mpc as rm
    b = synth.make_batch(200, shard_index=3)
    want = np.stack([rm.getFootPositionWorld(b["x_fb"][i], b["q"][i], rm.BipedParams()).reshape(6) for i in range(200)])
    np.testing.assert_allclose(b["pf_w"], want, rtol=0, atol=1e-13)


def test_simulator_adapter_clock_and_contact_rows_match_the_oracle_rule():
    """The adapter's own clock (biped_mpc_py_b200/sim.py): integer ticks, gait phase = tick % 10, contact rows of the table at
    MPC.py:52-55 - the rule R2 of oracle/rollout.py - checked without a GPU on the expression the adapter uses."""
    from oracle import rollout as ro
    import inspect
    from biped_mpc_py_b200 import sim
    src = inspect.getsource(sim.SimulatorAdapter.step)
    assert "rows < period // 2" in src and "self.tick % period" in src   # the expression mirrored below
    period, h = 10, 10
    for tick in range(0, 25):
        for gait in (0, 1):
            phase = np.array([tick % period])
            rows = (phase[:, None] + np.arange(h)[None, :]) % period
            left = rows < period // 2
            contact = np.where((np.array([gait]) == 1)[:, None, None], np.stack([left, ~left], axis=2), True).astype(np.uint8)[0]
            assert (contact == ro.contact_rows(tick, gait, h)).all(), (tick, gait)
