"""The oracle against (1) fixtures produced by the REAL reference module
(oracle/gen_golden.py) and (2) the KKT-certified known answers G1-G5 of SURVEY.md 8c."""
import numpy as np
import pytest

from oracle import qp_exact
from oracle import reference_mpc as rm
from conftest import variant_params


def test_contact_table_and_float_phase(golden):
    mpc = rm.MPCParams()
    for t, phase, contact in zip(golden["sweep_t"], golden["sweep_phase"], golden["sweep_contact"]):
        assert rm.gait_phase(t, mpc) == phase
        np.testing.assert_array_equal(rm.get_contact_sequence(t, mpc), contact)
        np.testing.assert_array_equal(rm.get_contact_sequence(t, mpc, extend=True), contact)
    # the quirk itself: (3*0.04)//0.04 == 2
    assert rm.gait_phase(3 * 0.04, mpc) == 2


def test_assembly_helpers_match_reference(golden):
    g = golden
    for c in range(len(g["t"])):
        mpc, biped = variant_params(int(g["variant"][c]))
        x_fb, t, q, qd = g["x_fb"][c], float(g["t"][c]), g["q"][c], g["qd"][c]
        pf_w = rm.getFootPositionWorld(x_fb, q, biped)
        np.testing.assert_allclose(pf_w.reshape(-1), g["pf_w"][c], rtol=0, atol=2e-15)
        contact = g["contact"][c]
        x_ref = rm.get_reference_trajectory(x_fb, mpc)
        np.testing.assert_allclose(x_ref, g["x_ref"][c], rtol=0, atol=1e-15)
        foot_ref = rm.get_reference_foot_trajectory(x_fb, t, g["pf_w"][c], mpc, contact)
        np.testing.assert_allclose(foot_ref, g["foot_ref"][c], rtol=0, atol=1e-15)
        np.testing.assert_allclose(
            rm.get_reference_foot_trajectory(x_fb, t, g["pf_w"][c], mpc, contact, extend=True), g["foot_ref"][c],
            rtol=0, atol=1e-15)
        for col, (ka, kb) in ((0, ("A0", "B0")), (7, ("A1", "B1"))):
            a, b = rm.get_simplified_dynamics(mpc, biped, g["x_ref"][c][:, col], g["foot_ref"][c][:, col])
            np.testing.assert_allclose(a, g[ka][c], rtol=0, atol=1e-14)
            np.testing.assert_allclose(b, g[kb][c], rtol=1e-12, atol=1e-14)
        for leg, side in ((0, 1), (1, -1)):
            jm, jf = rm.getLegKinematics(*q[5 * leg:5 * leg + 5], side)
            np.testing.assert_allclose(jm, g["Jm"][c][leg], rtol=0, atol=1e-15)
            f_sw = rm.swingLegControl(x_fb, t, g["pf_w"][c][3 * leg:3 * leg + 3], g["vf_w"][c][leg], mpc, side)
            np.testing.assert_allclose(f_sw.reshape(-1), g["F_swing"][c][leg], rtol=1e-13, atol=1e-12)


def test_dense_qp_matches_reference(golden):
    g = golden
    for slot, c in enumerate(g["dense_case_index"]):
        mpc, biped = variant_params(int(g["variant"][c]))
        qp = rm.build_qp(g["x_fb"][c], float(g["t"][c]), g["pf_w"][c], mpc, biped, g["contact"][c])
        for key in ("H", "f", "G", "hv", "A", "b"):
            np.testing.assert_allclose(qp[key], g["dense_" + key][slot], rtol=1e-13, atol=1e-14, err_msg=key)


def test_solve_and_torques_match_reference_pipeline(golden):
    g = golden
    for c in range(len(g["t"])):
        mpc, biped = variant_params(int(g["variant"][c]))
        states, controls, info = rm.solve_mpc(g["x_fb"][c], float(g["t"][c]), g["pf_w"][c], mpc, biped,
                                              g["contact"][c], return_info=True)
        scale = max(1.0, np.abs(g["controls"][c]).max())
        assert np.abs(controls - g["controls"][c]).max() <= 1e-8 * scale
        assert np.abs(states - g["states"][c]).max() <= 1e-9
        assert abs(info["obj"] - g["obj"][c]) <= 1e-9 * abs(g["obj"][c])
        tau = rm.lowLevelControl(g["x_fb"][c], float(g["t"][c]), g["pf_w"][c].reshape(6, 1), g["q"][c], g["qd"][c],
                                 mpc, biped, g["contact"][c], controls[0].reshape(-1, 1))
        assert np.abs(tau.reshape(-1) - g["tau"][c]).max() <= 1e-8


# ---- SURVEY.md 8c known answers (reference defaults, MPC.py:13-16) -------------------------------
G_CASES = {
    "G1": dict(t=0.0, gait=1, obj=-1915.4842356071,
               u0=[0.07083475514, 0, 120.3450920, 0, 0, 0, 0, -2.301143994, 0.006373774181, 0, 0, 0],
               tau=[-0.008888407989, 2.467074387, 2.326015988, -16.40630606, 2.303977384, -0.6425, 8.802408717,
                    -3.511269837, -1.712412665, -0.4]),
    "G2": dict(t=0.12, gait=1, obj=-2051.5069414455,
               u0=[0.02666335287, 0, 44.30726752, 0, 0, 0, 0, -0.8661879918, 0.002399192173, 0, 0, 0],
               tau=[-0.003345741199, 0.9082989842, 0.8755502144, -6.021190881, 0.8672545259, -0.6425, 9.539477517,
                    -3.511269837, 5.685070966, -0.4]),
    "G3": dict(t=0.28, gait=1, obj=-2095.7168072019,
               u0=[0, 0, 0, 0.08586539723, 23.15587180, 150.0206570, 0, 0, 0, 0, -2.950737588, 0.4554065019],
               tau=[0.6925, -9.785059105, -3.511269837, 5.685070966, -0.4, -0.140183338, -10.455971603, 2.980887246,
                    -20.370207527, 2.954172204]),
    "G4": dict(t=0.0, gait=0, obj=-2105.9928700880,
               u0=[0, 0, 80.53248922, 0, 0, 80.53248922, 0, -1.604486872, 0, 0, -1.604486872, 0],
               tau=[0, 1.650916029, 1.604486872, -10.92342836, 1.604486872, 0, -1.248253583, 1.604486872,
                    -10.92342836, 1.604486872]),
}


@pytest.mark.parametrize("name", sorted(G_CASES))
def test_known_answers(name):
    case = G_CASES[name]
    out = rm.mpc_tick(rm.X_FB0, case["t"], rm.Q0, rm.QD0, rm.MPCParams(), rm.BipedParams(), gait=case["gait"])
    assert abs(out["info"]["obj"] - case["obj"]) < 1e-8
    np.testing.assert_allclose(out["controls"][0], case["u0"], rtol=2e-9, atol=2e-9)
    np.testing.assert_allclose(out["tau"].reshape(-1), case["tau"], rtol=2e-9, atol=2e-9)
    np.testing.assert_allclose(out["pf_w"].reshape(-1),
                               [-0.02, 0.09, -0.003126983722081, -0.02, -0.09, -0.003126983722081], atol=1e-14)
    assert out["info"]["cert"]["stationarity"] < 1e-9


def test_known_answer_g1_stance_profile():
    out = rm.mpc_tick(rm.X_FB0, 0.0, rm.Q0, rm.QD0, rm.MPCParams(), rm.BipedParams(), gait=1)
    u = out["controls"]
    stance_fz = np.where(out["contact"][:, 0] == 1, u[:, 2], u[:, 5])
    np.testing.assert_allclose(stance_fz, [120.345092, 13.876829, 0, 0, 0, 500, 440.854842, 167.710656, 34.474761, 0],
                               atol=2e-6)


def test_known_answer_g5_h30_standing():
    out = rm.mpc_tick(rm.X_FB0, 0.0, rm.Q0, rm.QD0, rm.MPCParams(h=30), rm.BipedParams(), gait=0)
    u0 = out["controls"][0]
    np.testing.assert_allclose(u0[[2, 5]], 80.27621019, rtol=1e-9)
    np.testing.assert_allclose(u0[[7, 10]], -1.598819615, rtol=1e-8)


def test_h30_walking_raises_like_reference_without_extension():
    mpc = rm.MPCParams(h=30)
    with pytest.raises(Exception):
        rm.mpc_tick(rm.X_FB0, 0.0, rm.Q0, rm.QD0, mpc, rm.BipedParams(), gait=1, extend=False)
    out = rm.mpc_tick(rm.X_FB0, 0.0, rm.Q0, rm.QD0, mpc, rm.BipedParams(), gait=1, extend=True)
    assert out["controls"].shape == (30, 12)


def test_exact_solver_agrees_with_independent_fullsize_ipm(golden):
    """Second opinion: a different algorithm (full-size dense IPM at cvxopt-like tolerances)
    lands within its own stopping tolerance of the certified optimum."""
    g = golden
    for slot in (0, 1, 2):
        c = g["dense_case_index"][slot]
        args = [g["dense_" + k][slot] for k in ("H", "f", "G", "hv", "A", "b")]
        x_ipm, iters = qp_exact.solve_ipm_fullsize(*args)
        u_exact = g["controls"][c].reshape(-1)
        u_ipm = x_ipm[130:]
        assert iters < 60
        assert np.abs(u_ipm - u_exact).max() <= 5e-3 * max(1.0, np.abs(u_exact).max())


def test_third_party_solver_agrees_with_the_certified_optimum():
    """Independent second opinion: the HiGHS QP solver bundled with scipy, on the dense QP the reference hands
    to cvxopt.  HiGHS stops at its own tolerances, so: same objective to 1e-8 relative, never better than the certified
    optimum, forces within 0.2 N (the objective is flat in the forces: R = 1e-4)."""
    from oracle import highs_check as hc, qp_exact
    if not hc.available():
        pytest.skip("scipy's bundled HiGHS is not importable here")
    from biped_mpc_py_b200 import synth
    mpc, biped = rm.MPCParams(), rm.BipedParams()
    cases = []
    pf0 = rm.getFootPositionWorld(rm.X_FB0, rm.Q0, biped).reshape(-1)
    for c in G_CASES.values():
        contact = rm.get_contact_sequence(c["t"], mpc) if c["gait"] == 1 else np.ones((mpc.h, 2))
        cases.append((rm.X_FB0, c["t"], pf0, contact))
    b = synth.make_batch(6, shard_index=41)
    cases += [(b["x_fb"][i], float(b["t"][i]), b["foot"][i], b["contact"][i]) for i in range(6)]
    for x, t, foot, contact in cases:
        qp = rm.build_qp(x, t, foot, mpc, biped, contact)
        z, obj, status = hc.solve_qp(qp["H"], qp["f"], qp["G"], qp["hv"], qp["A"], qp["b"])
        ex = qp_exact.solve(qp["H"], qp["f"], qp["G"], qp["hv"], qp["A"], qp["b"])
        assert status == "Optimal"
        assert abs(obj - ex["obj"]) <= 1e-8 * max(1.0, abs(ex["obj"]))
        assert ex["obj"] <= obj + 1e-9 * max(1.0, abs(obj))          # the certified point is never worse
        assert np.abs(z[13 * mpc.h:] - ex["x"][13 * mpc.h:]).max() <= 0.2
