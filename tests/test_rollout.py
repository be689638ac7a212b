"""Closed-loop rollout (BASELINE.json configs[4]): the CUDA loop ``bmpc_rollout`` against the
CPU loop in ``oracle/rollout.py`` (same rules R1-R6), tick by tick and along whole trajectories, cold and
warm-started.  CPU-only tests pin the oracle loop's own consistency."""
import numpy as np
import pytest

X_ATOL = 1e-6     # state after <= 30 closed-loop ticks (solutions agree to ~1e-9; the loop is contractive)
U_RTOL = 1e-4     # north_star force/moment tolerance, per tick
TAU_ATOL = 1e-4   # north_star torque tolerance [N*m], per tick


def _params():
    from oracle import reference_mpc as rm
    from biped_mpc_py_b200 import synth
    mpc = rm.MPCParams()
    biped = rm.BipedParams()
    biped.f_min = synth.rollout_biped().f_min
    return mpc, biped


# ------------------------------------------------------------------------------------ CPU (oracle) tests
def test_oracle_plant_step_is_the_mpc_prediction():
    """R5: x+ = A_0[x;1] + B_0 u_0 equals the first predicted state X_0 of the same solve (MPC.py:206-208)."""
    from oracle import rollout as ro, reference_mpc as rm
    from biped_mpc_py_b200 import synth
    mpc, biped = _params()
    b = synth.make_rollout_batch(3, shard_index=1)
    for i in range(3):
        out = ro.tick_once(b["x"][i], b["foot"][i], int(b["tick"][i]), int(b["gait"][i]), b["q"][i], b["qd"][i], mpc, biped)
        xn = ro.srb_step(b["x"][i], b["foot"][i], out["controls"][0], mpc, biped)
        np.testing.assert_allclose(xn, out["states"][0][:12], atol=1e-9)


def test_oracle_contact_rows_match_reference_table():
    """R2 equals get_contact_sequence (MPC.py:50-59) whenever the float phase is exact."""
    from oracle import rollout as ro, reference_mpc as rm
    mpc = rm.MPCParams()
    for tick in range(25):
        ref = rm.get_contact_sequence((tick + 0.5) * mpc.dt, mpc)
        np.testing.assert_array_equal(ro.contact_rows(tick, 1, mpc.h), ref)
    assert (ro.contact_rows(3, 0, 10) == 1).all()


def test_oracle_closed_loop_settles_and_touchdown_rule():
    """Standing converges to the command; walking places the landing foot on the swing target (R6)."""
    from oracle import rollout as ro
    from biped_mpc_py_b200 import synth
    mpc, biped = _params()
    b = synth.make_rollout_batch(8, shard_index=2)
    i = int(np.nonzero(b["gait"] == 0)[0][0]) if (b["gait"] == 0).any() else 0
    r = ro.rollout(b["x"][i], b["foot"][i], int(b["tick"][i]), 0, b["q"][i], b["qd"][i], 40, mpc, biped)
    assert abs(r["x"][-1][5] - mpc.x_cmd[5]) < 5e-3 and np.abs(r["x"][-1][6:12]).max() < 5e-2
    np.testing.assert_array_equal(r["foot"][0], r["foot"][-1])  # standing: feet never move
    j = int(np.nonzero(b["gait"] == 1)[0][0])
    t0 = int(b["tick"][j])
    r = ro.rollout(b["x"][j], b["foot"][j], t0, 1, b["q"][j], b["qd"][j], 12, mpc, biped)
    moved = [k for k in range(12) if not np.array_equal(r["foot"][k], r["foot"][k + 1])]
    assert moved == [k for k in range(12) if (t0 + k + 1) % 5 == 0]  # one touchdown every 5 ticks
    for k in moved:
        leg = 0 if (t0 + k + 1) % 10 == 0 else 1
        side = 1.0 if leg == 0 else -1.0
        np.testing.assert_allclose(r["foot"][k + 1][3 * leg:3 * leg + 3], ro.touchdown_target(r["x"][k + 1], mpc, side))


def test_oracle_fall_reset_rule():
    """R7: a robot past the thresholds is put back on the reference's initial state with its own feet."""
    from oracle import rollout as ro, reference_mpc as rm
    mpc, biped = _params()
    x = rm.X_FB0.copy()
    x[1], x[7] = 0.79, 5.0
    foot = rm.getFootPositionWorld(x, rm.Q0, biped).reshape(-1)
    xn, fn, t = ro.advance(x, foot, 3, 0, np.zeros(12), mpc, biped, rm.Q0)
    np.testing.assert_array_equal(xn, rm.X_FB0)
    np.testing.assert_allclose(fn, [-0.02, 0.09, -0.003126983722081, -0.02, -0.09, -0.003126983722081], atol=1e-12)
    assert t == 4 and not ro.fallen(rm.X_FB0) and ro.fallen(np.full(12, np.nan))


# ------------------------------------------------------------------------------------ GPU tests
@pytest.mark.gpu
def test_rollout_fall_reset_matches_oracle():
    """R7 on the GPU: robots pushed over the thresholds are reset exactly like the CPU loop resets them."""
    import torch
    from oracle import rollout as ro
    from biped_mpc_py_b200 import BatchedMPC, MPC, synth
    mpc, biped = _params()
    b = synth.make_rollout_batch(4, shard_index=9)
    b["x"][0, 1], b["x"][0, 7] = 0.79, 5.0      # pitches over
    b["x"][1, 5], b["x"][1, 11] = 0.26, -2.0    # drops below z = 0.25
    b["x"][2, 6] = np.inf                       # non-finite
    s = BatchedMPC(MPC(), synth.rollout_biped(), max_batch=4)
    tn = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=s.device)
    st = [tn(b["x"], torch.float64), tn(b["foot"], torch.float64), tn(b["tick"], torch.int32), tn(b["gait"], torch.uint8),
          tn(b["q"], torch.float64), tn(b["qd"], torch.float64)]
    out = s.rollout(*st, 1, warm_start=False, n_log=4)
    torch.cuda.synchronize()
    sn = BatchedMPC.rollout_stats(out["stats"])
    assert sn["falls"] == 3 and sn["bad_input"] == 1, sn
    u0 = out["u0_log"].cpu().numpy()[0]
    for i in range(4):
        xn, fn, _ = ro.advance(b["x"][i], b["foot"][i], int(b["tick"][i]), int(b["gait"][i]), u0[i], mpc, biped, b["q"][i])
        np.testing.assert_allclose(st[0][i].cpu().numpy(), xn, atol=1e-12)
        np.testing.assert_allclose(st[1][i].cpu().numpy(), fn, atol=1e-12)
    s.close()


def _gpu_rollout(n, ticks, warm, n_log, shard=5, max_batch=None):
    import torch
    from biped_mpc_py_b200 import BatchedMPC, MPC, synth
    assert torch.cuda.is_available()
    b = synth.make_rollout_batch(n, shard_index=shard)
    s = BatchedMPC(MPC(), synth.rollout_biped(), max_batch=max_batch or n)
    dev = s.device
    tn = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    st = dict(x=tn(b["x"], torch.float64), foot=tn(b["foot"], torch.float64), tick=tn(b["tick"], torch.int32),
              gait=tn(b["gait"], torch.uint8), q=tn(b["q"], torch.float64), qd=tn(b["qd"], torch.float64))
    out = s.rollout(st["x"], st["foot"], st["tick"], st["gait"], st["q"], st["qd"], ticks, warm_start=warm, n_log=n_log)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    res["final_x"], res["final_foot"], res["final_tick"] = st["x"].cpu().numpy(), st["foot"].cpu().numpy(), st["tick"].cpu().numpy()
    res["stats_named"] = BatchedMPC.rollout_stats(out["stats"])
    s.close()
    return b, res


@pytest.mark.gpu
@pytest.mark.parametrize("warm", [False, True])
def test_rollout_matches_oracle_loop(warm):
    """Whole trajectories of 6 robots x 25 ticks against the CPU loop, and every logged tick re-solved by the oracle
    FROM THE GPU'S OWN STATE (no accumulation): forces, torques, next state, foothold."""
    from oracle import rollout as ro
    mpc, biped = _params()
    n, ticks = 6, 25
    b, res = _gpu_rollout(n, ticks, warm, n_log=n)
    assert res["stats_named"]["not_optimal"] == 0 and res["stats_named"]["robot_ticks"] == n * ticks
    np.testing.assert_array_equal(res["final_tick"], b["tick"] + ticks)
    np.testing.assert_array_equal(res["x_log"][-1], res["final_x"])
    for i in range(n):
        r = ro.rollout(b["x"][i], b["foot"][i], int(b["tick"][i]), int(b["gait"][i]), b["q"][i], b["qd"][i], ticks, mpc, biped)
        np.testing.assert_allclose(res["x_log"][:, i], r["x"], atol=X_ATOL)
        np.testing.assert_allclose(res["foot_log"][:, i], r["foot"], atol=X_ATOL)
        scale = np.maximum(1.0, np.abs(r["u0"]).max(axis=1, keepdims=True))
        assert (np.abs(res["u0_log"][:, i] - r["u0"]) / scale).max() <= U_RTOL
        assert np.abs(res["tau_log"][:, i] - r["tau"]).max() <= TAU_ATOL
    # tick-by-tick from the GPU's own logged state (every 4th tick of two robots keeps this in seconds)
    for i in (0, 1):
        for k in range(0, ticks, 4):
            x, foot, tick = res["x_log"][k, i], res["foot_log"][k, i], int(b["tick"][i]) + k
            out = ro.tick_once(x, foot, tick, int(b["gait"][i]), b["q"][i], b["qd"][i], mpc, biped)
            u0 = out["controls"][0]
            assert np.abs(res["u0_log"][k, i] - u0).max() / max(1.0, np.abs(u0).max()) <= U_RTOL
            assert np.abs(res["tau_log"][k, i] - out["tau"]).max() <= TAU_ATOL
            xn, fn, _ = ro.advance(x, foot, tick, int(b["gait"][i]), res["u0_log"][k, i], mpc, biped, b["q"][i])
            np.testing.assert_allclose(res["x_log"][k + 1, i], xn, atol=1e-12)
            np.testing.assert_allclose(res["foot_log"][k + 1, i], fn, atol=1e-12)


@pytest.mark.gpu
def test_rollout_warm_equals_cold_and_split_invariant():
    """Warm start changes the path to the optimum, not the optimum: same trajectories as the cold loop to 1e-9;
    and a rollout in two calls (40 = 15 + 25 ticks) equals one call."""
    n, ticks = 512, 40
    _, cold = _gpu_rollout(n, ticks, False, n_log=64)
    _, warm = _gpu_rollout(n, ticks, True, n_log=64)
    assert cold["stats_named"]["not_optimal"] == 0 and warm["stats_named"]["not_optimal"] == 0
    np.testing.assert_allclose(warm["final_x"], cold["final_x"], atol=1e-9)
    np.testing.assert_allclose(warm["u0_log"], cold["u0_log"], rtol=1e-7, atol=1e-7)
    assert warm["stats_named"]["warm_hit_rate"] > 0.5, warm["stats_named"]
    assert warm["stats_named"]["mean_iters"] < 0.5 * cold["stats_named"]["mean_iters"]
    # split invariance (cold, so the two runs take identical solver paths: bit-exact)
    import torch
    from biped_mpc_py_b200 import BatchedMPC, MPC, synth
    b = synth.make_rollout_batch(n, shard_index=5)
    s = BatchedMPC(MPC(), synth.rollout_biped(), max_batch=n)
    tn = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=s.device)
    st = [tn(b["x"], torch.float64), tn(b["foot"], torch.float64), tn(b["tick"], torch.int32), tn(b["gait"], torch.uint8),
          tn(b["q"], torch.float64), tn(b["qd"], torch.float64)]
    s.rollout(*st, 15, warm_start=False)
    s.rollout(*st, 25, warm_start=False)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(st[0].cpu().numpy(), cold["final_x"])
    s.close()


@pytest.mark.gpu
def test_rollout_long_horizon_stays_certified():
    """1,000 ticks x 2,048 robots (config 4's tick count on a subset): every tick certified optimal, states bounded,
    and the plant step equals the solver's own first predicted state on a spot-checked tick."""
    n, ticks = 2048, 1000
    _, res = _gpu_rollout(n, ticks, True, n_log=4)
    sn = res["stats_named"]
    assert sn["robot_ticks"] == n * ticks and sn["not_optimal"] == 0 and sn["bad_input"] == 0, sn
    assert np.isfinite(res["final_x"]).all()
    assert np.abs(res["final_x"][:, 0:3]).max() <= 0.8 and np.abs(res["final_x"][:, 5] - 0.55).max() < 0.3
    assert sn["falls"] < 0.001 * n * ticks / 100, sn  # R7 resets are rare (about 1 robot in 4,000 per 300 ticks)


@pytest.mark.gpu
def test_handle_warm_start_in_a_caller_owned_loop():
    """``bmpc_warm_start``: a loop the CALLER owns (one ``step`` per tick, states fed back by the caller - here the logged
    states of a device rollout) gives the same results warm and cold, and the warm ticks skip the interior point."""
    import torch
    from biped_mpc_py_b200 import BatchedMPC, MPC, synth
    n, ticks = 64, 12
    b, res = _gpu_rollout(n, ticks, False, n_log=n, shard=6)
    s = BatchedMPC(MPC(), synth.rollout_biped(), max_batch=n)
    dev = s.device
    tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    q, qd = tn(b["q"]), tn(b["qd"])
    outs = {}
    for warm in (False, True):
        s.warm_start(warm)
        u0, iters = [], []
        for k in range(ticks):
            tick = b["tick"] + k
            phase = tn((tick % 10).astype(np.int32), torch.int32)
            walking = b["gait"][:, None, None] == 1
            rows = (tick[:, None] + np.arange(10)[None, :]) % 10
            contact = np.where(walking, np.stack([rows < 5, rows >= 5], axis=2), True).astype(np.uint8)
            x, foot = tn(res["x_log"][k]), tn(res["foot_log"][k])
            out = s.step(x, phase, tn(tick * 0.04), foot, tn(contact, torch.uint8), q, qd, foot)
            torch.cuda.synchronize()
            assert (out["status"].cpu().numpy() == 0).all()
            u0.append(out["controls"][:, 0, :].cpu().numpy().copy())
            iters.append(out["iters"].cpu().numpy().copy())
        outs[warm] = (np.array(u0), np.array(iters))
    np.testing.assert_allclose(outs[True][0], outs[False][0], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(outs[False][0], res["u0_log"], rtol=1e-9, atol=1e-9)   # the caller's loop == bmpc_rollout
    assert outs[True][1][0].min() > 0                      # first warm call is cold
    assert (outs[True][1][1:] == 0).mean() > 0.9           # later ticks: straight to the polish
    s.reset_warm_start()
    s.close()


@pytest.mark.gpu
def test_handle_warm_start_with_a_growing_batch():
    """The warm-start store of a handle only describes the robots of the last call that went through: a call with MORE robots
    must start cold (the store's tail was never written - it is initialised to "no guess" - and warm_n gates it), a call with
    the same or fewer robots starts warm; results equal the cold solve either way."""
    import torch
    from biped_mpc_py_b200 import BatchedMPC, MPC, synth
    n = 48
    mpc, biped = MPC(), synth.rollout_biped()
    b = synth.make_batch(n, shard_index=31, mpc=mpc, biped=biped)
    keys = ("x_fb", "t", "foot", "contact", "q", "qd", "pf_w")
    cold = BatchedMPC(mpc, biped, max_batch=n)
    ref = cold.step_host(*[b[k] for k in keys])
    cold.close()
    s = BatchedMPC(mpc, biped, max_batch=n)
    s.warm_start(True)
    first = s.step_host(*[b[k][:16] for k in keys])          # cold (nothing stored), stores 16 robots
    assert first["iters"].min() > 0
    # warm: the same 16 robots from their own active sets shifted by one stage (these are unrelated random states, not a closed loop,
    # so only some of the guesses certify; those robots skip the interior point: 0 iterations)
    again = s.step_host(*[b[k][:16] for k in keys])
    assert (again["iters"] == 0).any()
    grown = s.step_host(*[b[k] for k in keys])               # 48 > 16: cold for everybody
    assert grown["iters"].min() > 0
    shrunk = s.step_host(*[b[k][:32] for k in keys])         # 32 <= 48: warm
    assert (shrunk["iters"] == 0).any()
    for out, m in ((first, 16), (again, 16), (grown, n), (shrunk, 32)):
        assert (out["status"] == 0).all()
        np.testing.assert_allclose(out["controls"], ref["controls"][:m], rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(out["tau"], ref["tau"][:m], rtol=1e-7, atol=1e-7)
    s.close()


@pytest.mark.gpu
def test_simulator_adapter_step_observe_loop_against_oracle():
    """SimulatorAdapter (biped_mpc_py_b200/sim.py: the reference main script MPC.py:475-495 as a step/observe object) in a loop
    with a CPU "simulator" - the oracle's plant step, joint angles held, feet by forward kinematics as MPC.py:478-479 - :
    at every control period the adapter's controls and torques equal the oracle tick on the same observation, warm-started
    or cold, and reset() restarts the clocks."""
    import biped_mpc_py_b200 as bm
    from biped_mpc_py_b200 import synth
    from oracle import reference_mpc as rm
    from oracle import rollout as ro
    mpc, biped = bm.MPC(), synth.rollout_biped()
    ompc, obiped = _params()
    n, ticks = 3, 7
    gait = np.array([1, 0, 1])
    b = synth.make_batch(n, shard_index=21, mpc=mpc, biped=biped)
    for warm in (True, False):
        ctl = bm.SimulatorAdapter(n, mpc, biped, gait=gait, warm_start=warm, strict=True)
        x = np.tile(rm.X_FB0, (n, 1)) + 0.02 * (b["x_fb"] - b["x_fb"].mean(axis=0))
        q, qd = b["q"].copy(), 0.1 * b["qd"]
        for k in range(ticks):
            ctl.observe(x, q, qd)
            tau = ctl.step()
            assert (ctl.last["status"] == 0).all() and (ctl.tick == k + 1).all()
            for i in range(n):
                pf = rm.getFootPositionWorld(x[i], q[i], obiped).reshape(6)
                np.testing.assert_allclose(ctl.last["pf_w"][i], pf, rtol=0, atol=1e-13)
                want = ro.tick_once(x[i], pf, k, int(gait[i]), q[i], qd[i], ompc, obiped)
                assert (ctl.last["contact"][i] == want["contact"]).all()
                scale = max(1.0, np.abs(want["controls"]).max())
                assert np.abs(ctl.last["controls"][i] - want["controls"]).max() / scale <= 1e-5, (warm, k, i)
                assert np.abs(tau[i] - want["tau"]).max() <= 1e-4, (warm, k, i)
            # the "simulator": the reference's own discretised model driven by the forces the controller asked for
            x = np.stack([ro.srb_step(x[i], ctl.last["pf_w"][i], ctl.last["controls"][i][0], ompc, obiped) for i in range(n)])
        ctl.reset()
        assert (ctl.tick == 0).all()
        ctl.close()
