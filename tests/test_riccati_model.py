"""CPU check of the stage-wise (Riccati) backend's math (tools/riccati_model.py, the model csrc/bmpc_riccati.cuh was
written from) against the dense condensed form of tools/kernel_model.py, itself pinned to the oracle by the GPU
assembly test: products Hc v + g, diag(Hc), and solves with barrier-like block-diagonal terms."""
import numpy as np
import pytest


@pytest.mark.parametrize("h", [10, 30])
def test_stagewise_backend_equals_dense_form(h):
    from oracle import reference_mpc as rm
    from tools import kernel_model as km
    from tools.riccati_model import Lqr
    from biped_mpc_py_b200 import synth, MPC
    rng = np.random.default_rng(h)
    mpc, biped = rm.MPCParams(h=h), rm.BipedParams()
    b = synth.make_batch(4, shard_index=21, mpc=MPC(h=h), extend=True, walking_prob=0.5)
    assert set(b["gait"].tolist()) == {0, 1}
    for i in range(4):
        red = km.build_reduced(b["x_fb"][i], int(b["phase_k"][i]), b["foot"][i], b["contact"][i], mpc, biped, extend=True)
        lq = Lqr(red, mpc, biped)
        n, LB = len(red["g"]), red["LB"]
        v = rng.normal(size=n) * 50
        ref = red["Hc"] @ v + red["g"]
        np.testing.assert_allclose(lq.grad(v), ref, rtol=0, atol=1e-12 * np.abs(ref).max())
        np.testing.assert_allclose(lq.hdiag(), np.diag(red["Hc"]), rtol=1e-13)
        Rt, M = [], red["Hc"].copy()
        for j in range(lq.nb):
            Cj = rng.normal(size=(11, LB))
            D = 10.0 ** rng.uniform(-6, 4, 11)        # barrier weights up to the hand-over point of the interior point
            blk = np.diag(lq.Rb[j]) + Cj.T @ (D[:, None] * Cj)
            Rt.append(blk)
            M[LB * j:LB * j + LB, LB * j:LB * j + LB] += blk - np.diag(lq.Rb[j])
        assert lq.factor(Rt)
        rhs = rng.normal(size=n) * 1e3
        x = lq.solve(rhs)
        assert np.abs(M @ x - rhs).max() <= 1e-8 * np.abs(rhs).max()
