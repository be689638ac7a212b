"""BASELINE.json's full single-GPU size (262,144 instances) through size-independent properties: every instance certified,
every returned control primal-feasible for the reference's constraint rows (MPC.py:220-271), swing feet and pinned
components exactly at their values, and the first predicted state equal to one step of the reference's own discretised
dynamics (MPC.py:148-185, 206-208) applied to the returned first-stage input."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rot_dyn(e):  # Rx(roll) Ry(pitch) Rz(yaw) with yaw = e0, pitch = e1, roll = e2 (MPC.py:151-156), vectorised
    cz, sz, cy, sy, cx, sx = np.cos(e[:, 0]), np.sin(e[:, 0]), np.cos(e[:, 1]), np.sin(e[:, 1]), np.cos(e[:, 2]), np.sin(e[:, 2])
    R = np.empty((len(e), 3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = cy * cz, -cy * sz, sy
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = -cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy
    return R


def test_262144_instances_properties():
    from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
    from biped_mpc_py_b200.synth import _eul2rotm_batch
    mpc, biped = MPC(), Biped()
    n, h = 262144, 10
    b = synth.make_batch(n, shard_index=3, mpc=mpc, biped=biped)
    s = BatchedMPC(mpc, biped, max_batch=n)
    out = s.step_host(b["x_fb"], b["t"], b["foot"], b["contact"], b["q"], b["qd"], b["pf_w"], want_states=True)
    s.close()
    assert (out["status"] == 0).all(), np.bincount(out["status"], minlength=4)
    U, X, c = out["controls"], out["states"], b["contact"].astype(bool)          # (n,h,12), (n,h,13), (n,h,2)
    assert np.isfinite(U).all() and np.isfinite(X).all() and (X[:, :, 12] == 1.0).all()
    scale = np.maximum(1.0, np.abs(U).reshape(n, -1).max(axis=1))[:, None]
    tol = 1e-9 * scale
    mu, fmax, tmax = float(biped.mu), np.asarray(biped.f_max).reshape(3), np.asarray(biped.tau_max).reshape(3)
    fmin, tmin = np.asarray(biped.f_min).reshape(3), np.asarray(biped.tau_min).reshape(3)
    R = _eul2rotm_batch(b["x_fb"][:, 0:3])                                      # current orientation, every stage (MPC.py:253)
    ez, ey = R[:, :, 2], R[:, :, 1]
    lt, lh = biped.lt - 0.01, biped.lh - 0.02                                   # MPC.py:254-255
    for leg in range(2):
        f, m = U[:, :, 3 * leg:3 * leg + 3], U[:, :, 6 + 3 * leg:9 + 3 * leg]
        on = c[:, :, leg]
        assert (f[~on] == 0).all() and (m[~on] == 0).all()                       # swing foot: box scaled by contact (MPC.py:234-251)
        assert (m[:, :, 0] == 0).all()                                          # tau_max[0] = 0 pins mx (MPC.py:47)
        for a in range(2):                                                      # friction pyramid, MPC.py:220-232
            assert ((f[:, :, a] - mu * f[:, :, 2]) <= tol).all() and ((-f[:, :, a] - mu * f[:, :, 2]) <= tol).all()
        for a in range(3):
            assert (f[:, :, a] <= fmax[a] * on + tol).all() and (f[:, :, a] >= fmin[a] * on - tol).all()
            assert (m[:, :, a] <= tmax[a] * on + tol).all() and (m[:, :, a] >= tmin[a] * on - tol).all()
        fz_b = np.einsum("nc,nsc->ns", ez, f)                                   # line foot, MPC.py:253-271
        my_b = np.einsum("nc,nsc->ns", ey, m)
        assert ((-lh * fz_b + my_b) <= tol).all() and ((-lt * fz_b - my_b) <= tol).all()
    # first predicted state = A_0 [x;1] + B_0 u_0 (MPC.py:148-185, 206-208), vectorised
    x, dt = b["x_fb"], float(mpc.dt)
    rot = _rot_dyn(x[:, 0:3])
    iw = np.einsum("nji,jk,nkl->nil", rot, np.asarray(biped.I, dtype=float), rot)
    cz, sz, cp, sp = np.cos(x[:, 0]), np.sin(x[:, 0]), np.cos(x[:, 1]), np.sin(x[:, 1])
    M = np.zeros((n, 3, 3))
    M[:, 0, 0], M[:, 0, 1], M[:, 1, 0], M[:, 1, 1], M[:, 2, 0], M[:, 2, 2] = cz * cp, -sz, sz * cp, cz, -sp, 1.0
    rinv = np.linalg.inv(M)
    u0 = U[:, 0, :]
    mom = np.zeros((n, 3))
    frc = np.zeros((n, 3))
    for leg in range(2):
        r = b["foot"][:, 3 * leg:3 * leg + 3] - x[:, 3:6]
        mom += np.cross(r, u0[:, 3 * leg:3 * leg + 3]) + u0[:, 6 + 3 * leg:9 + 3 * leg]
        frc += u0[:, 3 * leg:3 * leg + 3]
    xn = x.copy()
    xn[:, 0:3] += dt * np.einsum("nij,nj->ni", rinv, x[:, 6:9])
    xn[:, 3:6] += dt * x[:, 9:12]
    xn[:, 6:9] += dt * np.linalg.solve(iw, mom[:, :, None])[:, :, 0]
    xn[:, 9:12] += dt * frc / float(biped.m)
    xn[:, 11] -= dt * float(biped.g)
    err = np.abs(X[:, 0, :12] - xn).max(axis=1) / np.maximum(1.0, np.abs(xn).max(axis=1))
    assert err.max() <= 1e-10, err.max()
