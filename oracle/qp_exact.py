"""Exact solve of the reference's QP (stands in for ``cvxopt.solvers.qp``, MPC.py:289-297).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

    min 1/2 x'Hx + f'x   s.t.  G x <= hv,  A x = b

cvxopt (unpinned, not installable here) is a dense primal-dual interior-point
method; its published algorithm stops at abstol=1e-7 / reltol=1e-6 / feastol=1e-7,
which leaves it ~1e-4 relative away from the optimum on this problem (SURVEY.md
7.1).  The QP is strictly convex, so the optimum is unique; this module returns
*that* point, independent of any solver's stopping rule:

  1. presolve    - variables pinned by opposing bound rows (swing-foot forces,
                   tau_max[0] = 0, MPC.py:47, 241-248) are fixed; without this the
                   feasible set has no interior and interior-point methods stall.
  2. null space  - equalities removed with an orthonormal basis from an SVD of A
                   (generic linear algebra: no knowledge of the dynamics structure,
                   so this is an independent check of the product's condensing).
  3. interior point (Mehrotra predictor-corrector) on the reduced dense problem.
  4. polish      - rows with small slack become equalities, null-space solve,
                   multipliers by NNLS, active set repaired until consistent.
  5. certificate - KKT residuals of the ORIGINAL problem (stationarity with
                   multipliers >= 0, primal feasibility, complementarity).

``solve_ipm_fullsize`` is the "cvxopt-class stand-in" used only as the timed CPU
baseline: the same presolve, then a dense interior-point iteration on the full
(states + inputs, equality-constrained) system stopped at cvxopt's default-like
tolerances, i.e. the per-iteration dense factorisation the reference pays for.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_factor, cho_solve, lu_factor, lu_solve
from scipy.optimize import nnls


class QPError(RuntimeError):
    pass


# ----------------------------------------------------------------------------
# presolve
# ----------------------------------------------------------------------------

def _presolve_bounds(G, hv):
    """Bounds implied by single-entry rows of G; returns lb, ub and the row of each."""
    n = G.shape[1]
    lb = np.full(n, -np.inf)
    ub = np.full(n, np.inf)
    lb_row = np.full(n, -1)
    ub_row = np.full(n, -1)
    nnz = np.count_nonzero(G, axis=1)
    for r in np.nonzero(nnz == 1)[0]:
        j = int(np.nonzero(G[r])[0][0])
        a = G[r, j]
        bound = hv[r] / a
        if a > 0:
            if bound < ub[j]:
                ub[j], ub_row[j] = bound, r
        else:
            if bound > lb[j]:
                lb[j], lb_row[j] = bound, r
    return lb, ub, lb_row, ub_row


# ----------------------------------------------------------------------------
# dense Mehrotra interior point for  min 1/2 w'Pw + q'w,  G w <= h
# ----------------------------------------------------------------------------

def _ipm_ineq(P, q, G, h, tol=1e-10, maxit=100):
    n, m = len(q), len(h)
    eye = np.eye(n)
    w = np.linalg.solve(P + 1e-9 * np.trace(P) / n * eye, -q)
    s = np.maximum(h - G @ w, 1.0)
    z = np.ones(m)
    scale_d = 1.0 + np.abs(q).max()
    scale_p = 1.0 + np.abs(h).max()
    it = 0
    for it in range(1, maxit + 1):
        rd = P @ w + q + G.T @ z
        rp = G @ w + s - h
        mu = float(s @ z) / m
        if np.abs(rd).max() <= tol * scale_d and np.abs(rp).max() <= tol * scale_p and mu <= tol:
            break
        d = z / s
        M = P + G.T @ (d[:, None] * G)
        try:
            cf = cho_factor(M, lower=True, check_finite=False)
        except np.linalg.LinAlgError:
            cf = cho_factor(M + 1e-12 * np.trace(M) / n * eye, lower=True, check_finite=False)

        def newton(rc):
            dw = cho_solve(cf, -rd - G.T @ (d * rp - rc / s), check_finite=False)
            ds = -rp - G @ dw
            dz = (-rc - z * ds) / s
            return dw, ds, dz

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0

        dw_a, ds_a, dz_a = newton(s * z)
        a_aff = min(max_step(s, ds_a), max_step(z, dz_a))
        mu_aff = float((s + a_aff * ds_a) @ (z + a_aff * dz_a)) / m
        sigma = (mu_aff / mu) ** 3
        dw, ds, dz = newton(s * z + ds_a * dz_a - sigma * mu)
        alpha = min(1.0, 0.995 * min(max_step(s, ds), max_step(z, dz)))
        w = w + alpha * dw
        s = s + alpha * ds
        z = z + alpha * dz
    return w, s, z, it


# ----------------------------------------------------------------------------
# active-set polish in the reduced space
# ----------------------------------------------------------------------------

def _nullspace(M, rtol=1e-11):
    """Orthonormal null-space basis and pseudo-inverse pieces of a (possibly rank-deficient) M."""
    if M.shape[0] == 0:
        return np.eye(M.shape[1]), None
    U, S, Vt = np.linalg.svd(M, full_matrices=True)
    rank = int(np.sum(S > rtol * max(S.max(), 1e-300)))
    return Vt[rank:].T, (U[:, :rank], S[:rank], Vt[:rank])


def _eq_qp(P, q, Ga, ha):
    """min 1/2 w'Pw + q'w  s.t.  Ga w = ha  (Ga may have dependent rows)."""
    N, piv = _nullspace(Ga)
    if piv is None:
        wp = np.zeros(len(q))
        incons = 0.0
    else:
        U, S, Vt = piv
        wp = Vt.T @ ((U.T @ ha) / S)
        incons = float(np.abs(Ga @ wp - ha).max())
    if N.shape[1] == 0:
        return wp, incons
    Pn = N.T @ P @ N
    v = np.linalg.solve(Pn, -N.T @ (P @ wp + q))
    return wp + N @ v, incons


def _polish(P, q, G, h, w0, slack0, act_tol, feas_tol=1e-9):
    scale = 1.0 + np.abs(h).max()
    active = slack0 <= act_tol * scale
    m = len(h)
    w, lam, resid = w0, np.zeros(m), np.inf
    grad_scale = 1.0
    for _ in range(40):
        idx = np.nonzero(active)[0]
        w, incons = _eq_qp(P, q, G[idx], h[idx])
        if incons > 1e-8 * scale:
            return None
        viol = G @ w - h
        viol[idx] = 0.0
        worst = int(np.argmax(viol))
        if viol[worst] > feas_tol * scale:
            active[viol > feas_tol * scale] = True
            continue
        g = P @ w + q
        grad_scale = 1.0 + np.abs(q).max()
        lam = np.zeros(m)
        if len(idx):
            lam_a, resid = nnls(G[idx].T, -g, maxiter=50 * len(idx) + 200)
            lam[idx] = lam_a
        else:
            resid = float(np.linalg.norm(g))
        if resid <= 1e-9 * grad_scale:
            return w, lam, resid / grad_scale
        # drop the active row that least-squares wants most negative
        lam_ls = np.linalg.lstsq(G[idx].T, -g, rcond=None)[0]
        j = int(np.argmin(lam_ls))
        if lam_ls[j] >= 0:
            return None
        active[idx[j]] = False
    return None


# ----------------------------------------------------------------------------
# public entry
# ----------------------------------------------------------------------------

def solve(H, f, G, hv, A, b, full_certificate: bool = False, ipm_tol: float = 1e-10):
    """Unique optimum of the strictly convex QP with a KKT certificate.

    Returns dict(x, lam, nu, obj, iters, cert) where ``cert`` holds the residuals of
    the original problem: stationarity, primal_ineq, primal_eq, dual_min,
    complementarity (all absolute, inf-norm).  Raises ``QPError`` if no certified
    point is found.  (``full_certificate`` is accepted for API symmetry; the
    certificate is always evaluated on the original problem.)
    """
    H = np.asarray(H, dtype=float)
    f = np.asarray(f, dtype=float).reshape(-1)
    G = np.asarray(G, dtype=float)
    hv = np.asarray(hv, dtype=float).reshape(-1)
    A = np.asarray(A, dtype=float)
    b = np.asarray(b, dtype=float).reshape(-1)
    n = len(f)

    lb, ub, lb_row, ub_row = _presolve_bounds(G, hv)
    if np.any(ub < lb - 1e-12):
        raise QPError("infeasible bounds")
    fixed = (ub - lb) <= 1e-12
    free = ~fixed
    x_fix = np.where(fixed, lb, 0.0)

    A_f = A[:, free]
    b_f = b - A[:, fixed] @ x_fix[fixed]
    Z, piv = _nullspace(A_f)
    U, S, Vt = piv
    if len(S) < A.shape[0]:
        raise QPError("equality rows are rank deficient")
    x_p = Vt.T @ ((U.T @ b_f) / S)

    H_ff = H[np.ix_(free, free)]
    lin = f[free] + H[np.ix_(free, fixed)] @ x_fix[fixed]
    P = Z.T @ H_ff @ Z
    P = 0.5 * (P + P.T)
    q = Z.T @ (H_ff @ x_p + lin)
    G_f = G[:, free]
    Gr_all = G_f @ Z
    hr_all = hv - G[:, fixed] @ x_fix[fixed] - G_f @ x_p
    rownorm = np.abs(Gr_all).max(axis=1)
    keep = rownorm > 1e-13
    if np.any(hr_all[~keep] < -1e-9):
        raise QPError("infeasible after fixing pinned variables")
    Gr, hr = Gr_all[keep], hr_all[keep]

    w, s, z, iters = _ipm_ineq(P, q, Gr, hr, tol=ipm_tol)
    slack = hr - Gr @ w

    result = None
    for act_tol in (1e-6, 1e-5, 1e-7, 1e-4, 1e-8, 1e-3):
        result = _polish(P, q, Gr, hr, w, slack, act_tol)
        if result is not None:
            break
    if result is None:
        raise QPError("active-set polish did not certify")
    w, lam_r, _ = result

    x = np.zeros(n)
    x[free] = x_p + Z @ w
    x[fixed] = x_fix[fixed]

    # ---- multipliers and certificate of the ORIGINAL problem ----
    lam = np.zeros(len(hv))
    lam[np.nonzero(keep)[0]] = lam_r
    g_full = H @ x + f + G.T @ lam
    # nu from the free rows: g_full[free] lies in range(A_f') at the optimum
    nu = -(U @ ((Vt @ g_full[free]) / S))
    r = g_full + A.T @ nu
    for j in np.nonzero(fixed)[0]:
        # the pinned pair absorbs the residual with non-negative multipliers
        if r[j] < 0:
            lam[ub_row[j]] += -r[j] / G[ub_row[j], j]
        else:
            lam[lb_row[j]] += r[j] / -G[lb_row[j], j]
    r = H @ x + f + G.T @ lam + A.T @ nu
    slack_full = hv - G @ x
    cert = dict(
        stationarity=float(np.abs(r).max()),
        primal_ineq=float(max(0.0, -slack_full.min())),
        primal_eq=float(np.abs(A @ x - b).max()),
        dual_min=float(lam.min()),
        complementarity=float(np.abs(lam * slack_full).max()),
    )
    gscale = 1.0 + np.abs(f).max()
    if (cert["stationarity"] > 1e-8 * gscale or cert["primal_ineq"] > 1e-8 * (1 + np.abs(hv).max())
            or cert["primal_eq"] > 1e-8 * (1 + np.abs(b).max()) or cert["dual_min"] < 0
            or cert["complementarity"] > 1e-6 * gscale):
        raise QPError(f"KKT certificate failed: {cert}")
    obj = float(0.5 * x @ H @ x + f @ x)
    return dict(x=x, lam=lam, nu=nu, obj=obj, iters=iters, cert=cert)


# ----------------------------------------------------------------------------
# timed CPU baseline: "cvxopt-class stand-in"
# ----------------------------------------------------------------------------

def solve_ipm_fullsize(H, f, G, hv, A, b, abstol=1e-7, reltol=1e-6, feastol=1e-7, maxit=100):
    """Dense primal-dual interior point on the full equality-constrained system.

    Mirrors what the reference pays per call at MPC.py:297: every iteration
    factorises a dense KKT matrix of the full problem size (states + inputs +
    equalities).  Pinned variables are fixed first (see module docstring).  Stops
    at cvxopt's documented default tolerances; NOT the exact optimum.
    """
    H = np.asarray(H, dtype=float)
    f = np.asarray(f, dtype=float).reshape(-1)
    lb, ub, _, _ = _presolve_bounds(G, hv)
    fixed = (ub - lb) <= 1e-12
    free = ~fixed
    x_fix = np.where(fixed, lb, 0.0)
    Hf = H[np.ix_(free, free)]
    ff = f[free] + H[np.ix_(free, fixed)] @ x_fix[fixed]
    Af = A[:, free]
    bf = b - A[:, fixed] @ x_fix[fixed]
    Gf_all = G[:, free]
    hf_all = hv - G[:, fixed] @ x_fix[fixed]
    keep = np.abs(Gf_all).max(axis=1) > 0
    Gf, hf = Gf_all[keep], hf_all[keep]
    n, m, p = len(ff), len(hf), len(bf)

    x = np.zeros(n)
    y = np.zeros(p)
    s = np.maximum(hf - Gf @ x, 1.0)
    z = np.ones(m)
    K = np.zeros((n + p, n + p))
    K[:n, n:] = Af.T
    K[n:, :n] = Af
    it = 0
    for it in range(1, maxit + 1):
        rd = Hf @ x + ff + Gf.T @ z + Af.T @ y
        rp = Gf @ x + s - hf
        re = Af @ x - bf
        gap = float(s @ z)
        pcost = float(0.5 * x @ Hf @ x + ff @ x)
        pres = max(np.linalg.norm(rp), np.linalg.norm(re)) / max(1.0, np.linalg.norm(hf))
        dres = np.linalg.norm(rd) / max(1.0, np.linalg.norm(ff))
        if pres <= feastol and dres <= feastol and (gap <= abstol or gap / max(abs(pcost), 1e-300) <= reltol):
            break
        d = z / s
        K[:n, :n] = Hf + Gf.T @ (d[:, None] * Gf)
        lu = lu_factor(K, check_finite=False)  # dense factorisation of the full KKT matrix

        def newton(rc):
            rhs = np.concatenate([-rd - Gf.T @ (d * rp - rc / s), -re])
            sol = lu_solve(lu, rhs, check_finite=False)
            dx, dy = sol[:n], sol[n:]
            ds = -rp - Gf @ dx
            dz = (-rc - z * ds) / s
            return dx, dy, ds, dz

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0

        mu = gap / m
        dx, dy, ds, dz = newton(s * z)
        a = min(max_step(s, ds), max_step(z, dz))
        sigma = (float((s + a * ds) @ (z + a * dz)) / m / mu) ** 3
        dx, dy, ds, dz = newton(s * z + ds * dz - sigma * mu)
        a = min(1.0, 0.99 * min(max_step(s, ds), max_step(z, dz)))
        x, y, s, z = x + a * dx, y + a * dy, s + a * ds, z + a * dz
    full = np.zeros(len(f))
    full[free] = x
    full[fixed] = x_fix[fixed]
    return full, it
