"""Numpy model of the CUDA kernel's algorithm (design tool, not product, not oracle).

Builds the contact-reduced condensed QP the kernel solves and runs the same
ADMM -> active-set -> polish pipeline in numpy so parameters (rho, scaling, check
cadence) can be chosen and the CUDA code debugged stage by stage.

Reduced form.  Per stance foot-stage ("block") j = (stage s, foot l) the free
components of [f_l; m_l] (those whose box bounds differ) are the variables; swing
feet and pinned components are eliminated (they are forced to their bound).
States are condensed out with the structure of the reference's A_k (MPC.py:165-184):
    theta_i = theta_{i-1} + dt*Rinv_i*omega_{i-1},  p_i = p_{i-1} + dt*v_{i-1},
    omega_i = omega_{i-1} + W_i U_i,  v_i = v_{i-1} - dt*g*e3 + V_i U_i
so  dX_i/du_j = [dt*(P_i-P_s) W_j ; dt*(i-s) V_j ; W_j ; V_j],  P_i = sum_{l=1..i} Rinv_l.
Objective is the reference's divided by 2:  1/2 u'Hc u + g'u.
"""
from __future__ import annotations

import numpy as np

GAIT_PERIOD = 10


def _rot_dyn(yaw, pitch, roll):
    cz, sz, cy, sy, cx, sx = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    ry = np.array([[cy, 0, sy], [0, 1.0, 0], [-sy, 0, cy]])
    rx = np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return rx @ ry @ rz


def _rinv(yaw, pitch):
    # closed-form inverse of [[cy*cp,-sy,0],[sy*cp,cy,0],[-sp,0,1]]
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    return np.array([[cy / cp, sy / cp, 0.0], [-sy, cy, 0.0], [cy * sp / cp, sy * sp / cp, 1.0]])


def _eul2rotm(e):
    (cr, cp, cy), (sr, sp, sy) = np.cos(e), np.sin(e)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def build_reduced(x_fb, phase_k, foot, contact, mpc, biped, extend=False):
    """Return dict with Hc, g, blocks, per-block constraint data, and maps back to U."""
    h, dt = mpc.h, mpc.dt
    x_fb = np.asarray(x_fb, float)
    foot = np.asarray(foot, float).reshape(6)
    contact = np.asarray(contact).astype(int)
    Q = np.asarray(mpc.Q, float)
    R = np.asarray(mpc.R, float)
    x_cmd = np.asarray(mpc.x_cmd, float)
    # references
    x_ref = np.empty((12, h))
    x_ref[:, :] = x_cmd[:, None]
    x_ref[:, 0] = x_fb
    for i in range(6):
        if x_cmd[i + 6] != 0:
            for k in range(1, h):
                x_ref[i, k] = x_fb[i] + x_cmd[i + 6] * (k * dt)
    ex = mpc.kv * (x_fb[3] - x_cmd[3])
    ey = mpc.kv * (x_fb[4] - x_cmd[4])
    f1 = np.array([x_fb[3] + x_fb[9] * 1 / 2 * h / 2 * dt + ex, x_fb[4] + x_fb[10] * 1 / 2 * h / 2 * dt + ey, 0.0])
    f2 = np.array([x_fb[3] + x_fb[9] * 1 / 2 * h * dt + ex, x_fb[10] + x_fb[10] * 1 / 2 * h * dt + ey, 0.0])
    kk = phase_k % 5
    foot_ref = np.empty((6, h))
    if contact[0].sum() == 1:
        for k in range(h):
            if k < 5 - kk:
                foot_ref[:, k] = foot
            elif k < 10 - kk:
                foot_ref[:, k] = np.concatenate([f1, f1])
            else:
                foot_ref[:, k] = np.concatenate([f2, f2])
    else:
        foot_ref[:, :] = foot[:, None]

    # per-stage dynamics pieces
    Rinv = np.zeros((h, 3, 3))
    Iw_inv = np.zeros((h, 3, 3))
    for k in range(h):
        yaw, pitch, roll = x_ref[0, k], x_ref[1, k], x_ref[2, k]
        rot = _rot_dyn(yaw, pitch, roll)
        Iw_inv[k] = np.linalg.inv(rot.T @ np.asarray(biped.I, float) @ rot)
        Rinv[k] = _rinv(yaw, pitch)
    P = np.zeros((h, 3, 3))
    for k in range(1, h):
        P[k] = P[k - 1] + Rinv[k]

    # free response and error e_i = Xfree_i - xref_i
    th, p, om, v = x_fb[0:3].copy(), x_fb[3:6].copy(), x_fb[6:9].copy(), x_fb[9:12].copy()
    e = np.zeros((h, 12))
    gvec = np.array([0, 0, -biped.g])
    for i in range(h):
        th = th + dt * Rinv[i] @ om
        p = p + dt * v
        v = v + dt * gvec
        e[i] = np.concatenate([th, p, om, v]) - x_ref[:, i]

    # bounds per component [fx,fy,fz,mx,my,mz] when in contact
    lo6 = np.concatenate([np.asarray(biped.f_min, float).reshape(3), np.asarray(biped.tau_min, float).reshape(3)])
    hi6 = np.concatenate([np.asarray(biped.f_max, float).reshape(3), np.asarray(biped.tau_max, float).reshape(3)])
    comps = [c for c in range(6) if hi6[c] > lo6[c]]
    pinned = [c for c in range(6) if not hi6[c] > lo6[c]]
    LB = len(comps)
    rot_now = _eul2rotm(x_fb[0:3])
    ez, eyw = rot_now[:, 2], rot_now[:, 1]
    lt, lh = biped.lt - 0.01, biped.lh - 0.02
    mu = biped.mu

    blocks = [(s, l) for s in range(h) for l in range(2) if contact[s, l]]
    nb = len(blocks)
    n = LB * nb
    W = np.zeros((nb, 3, LB))
    V = np.zeros((nb, 3, LB))
    # pinned components contribute a constant input: fold into the free response
    Wp = np.zeros((nb, 3))
    Vp = np.zeros((nb, 3))
    for j, (s, l) in enumerate(blocks):
        r = foot_ref[3 * l:3 * l + 3, s] - x_ref[3:6, s]
        Bw = dt * Iw_inv[s] @ np.hstack([_skew(r), np.eye(3)])   # 3x6 for [f;m]
        Bv = dt / biped.m * np.hstack([np.eye(3), np.zeros((3, 3))])
        W[j] = Bw[:, comps]
        V[j] = Bv[:, comps]
        for c in pinned:
            Wp[j] += Bw[:, c] * lo6[c]
            Vp[j] += Bv[:, c] * lo6[c]
    # effect of pinned (non-zero) components on the error trajectory
    if pinned and np.any(lo6[pinned] != 0):
        for j, (s, l) in enumerate(blocks):
            for i in range(s, h):
                e[i, 0:3] += dt * (P[i] - P[s]) @ Wp[j]
                e[i, 3:6] += dt * (i - s) * Vp[j]
                e[i, 6:9] += Wp[j]
                e[i, 9:12] += Vp[j]

    Qth, Qp, Qw, Qv = Q[0:3], Q[3:6], Q[6:9], Q[9:12]
    Hc = np.zeros((n, n))
    g = np.zeros(n)
    for j, (s, l) in enumerate(blocks):
        gj = np.zeros(LB)
        for i in range(s, h):
            Gth = dt * (P[i] - P[s]) @ W[j]
            gj += Gth.T @ (Qth * e[i, 0:3]) + dt * (i - s) * V[j].T @ (Qp * e[i, 3:6]) \
                + W[j].T @ (Qw * e[i, 6:9]) + V[j].T @ (Qv * e[i, 9:12])
        g[LB * j:LB * j + LB] = gj
        for j2, (s2, l2) in enumerate(blocks):
            if j2 < j:
                continue
            blk = np.zeros((LB, LB))
            for i in range(max(s, s2), h):
                G1 = dt * (P[i] - P[s]) @ W[j]
                G2 = dt * (P[i] - P[s2]) @ W[j2]
                blk += G1.T @ (Qth[:, None] * G2) + dt * dt * (i - s) * (i - s2) * V[j].T @ (Qp[:, None] * V[j2]) \
                    + W[j].T @ (Qw[:, None] * W[j2]) + V[j].T @ (Qv[:, None] * V[j2])
            if j2 == j:
                Rb = np.concatenate([R[3 * l:3 * l + 3], R[6 + 3 * l:9 + 3 * l]])[comps]
                blk += np.diag(Rb)
                # pinned components' R-cost is constant; dropped
            Hc[LB * j:LB * j + LB, LB * j2:LB * j2 + LB] = blk
            Hc[LB * j2:LB * j2 + LB, LB * j:LB * j + LB] = blk.T

    # per-block constraint rows (same for every block): friction(4) + line-foot(2), in block coordinates
    F6 = np.zeros((6, 6))
    F6[0, [0, 2]] = [1, -mu]
    F6[1, [1, 2]] = [1, -mu]
    F6[2, [0, 2]] = [-1, -mu]
    F6[3, [1, 2]] = [-1, -mu]
    F6[4, 0:3], F6[4, 3:6] = -lh * ez, eyw
    F6[5, 0:3], F6[5, 3:6] = -lt * ez, -eyw
    Fb = F6[:, comps]
    fb_hi = -F6[:, pinned] @ lo6[pinned] if pinned else np.zeros(6)  # rows: F u <= fb_hi
    return dict(Hc=Hc, g=g, blocks=blocks, comps=comps, pinned=pinned, LB=LB, lo=lo6[comps], hi=hi6[comps],
                lo6=lo6, hi6=hi6, Fb=Fb, fb_hi=fb_hi, h=h, x_ref=x_ref, foot_ref=foot_ref, contact=contact,
                W=W, V=V, P=P, e=e)


def expand_controls(red, u):
    """Reduced vector -> controls (h,12) with swing feet 0 and pinned comps at their bound."""
    h = red["h"]
    U = np.zeros((h, 12))
    lo6 = red["lo6"]
    for j, (s, l) in enumerate(red["blocks"]):
        full = np.zeros(6)
        full[red["pinned"]] = lo6[red["pinned"]]
        full[red["comps"]] = u[red["LB"] * j:red["LB"] * j + red["LB"]]
        U[s, 3 * l:3 * l + 3] = full[0:3]
        U[s, 6 + 3 * l:9 + 3 * l] = full[3:6]
    return U


def constraint_matrix(red):
    """C = [I; blockdiag(Fb)], l, u for  l <= C u <= u."""
    nb, LB = len(red["blocks"]), red["LB"]
    n = nb * LB
    C = np.zeros((n + 6 * nb, n))
    C[:n, :n] = np.eye(n)
    lo = np.concatenate([np.tile(red["lo"], nb), np.full(6 * nb, -np.inf)])
    hi = np.concatenate([np.tile(red["hi"], nb), np.tile(red["fb_hi"], nb)])
    for j in range(nb):
        C[n + 6 * j:n + 6 * j + 6, LB * j:LB * j + LB] = red["Fb"]
    return C, lo, hi


# ----------------------------------------------------------------------------
# model of the kernel's interior-point solve
# ----------------------------------------------------------------------------

def block_rows(red, prune=True):
    """Per-block inequality rows  Cb u_b <= rb  in block coordinates (same for every block).

    Row order: lower bounds (LB), upper bounds (LB), friction (4), line-foot (2).
    With ``prune`` rows implied by others for these parameter values are dropped
    (host-side presolve; does not change the feasible set).
    """
    LB, comps, lo6, hi6 = red["LB"], red["comps"], red["lo6"], red["hi6"]
    rows, rhs, tags = [], [], []
    for k, c in enumerate(comps):
        e = np.zeros(LB); e[k] = -1.0
        rows.append(e); rhs.append(-lo6[c]); tags.append(("lo", c))
    for k, c in enumerate(comps):
        e = np.zeros(LB); e[k] = 1.0
        rows.append(e); rhs.append(hi6[c]); tags.append(("hi", c))
    for r in range(6):
        rows.append(red["Fb"][r].copy()); rhs.append(red["fb_hi"][r]); tags.append(("fric" if r < 4 else "line", r))
    rows, rhs = np.array(rows), np.array(rhs)
    keep = np.ones(len(rows), bool)
    if prune:
        mu = -red["Fb"][0][comps.index(2)] if 2 in comps else 0.0
        fx_free, fy_free, fz_free = 0 in comps, 1 in comps, 2 in comps
        if fx_free and fy_free and fz_free and mu > 0:
            # |fx| <= mu fz, |fy| <= mu fz  => fz >= 0, |fx|,|fy| <= mu*fz_max
            for i, (kind, c) in enumerate(tags):
                if kind == "lo" and c == 2 and lo6[2] <= 0:
                    keep[i] = False                      # fz >= lo (<=0) implied by the pyramid
                if kind == "hi" and c in (0, 1) and hi6[c] >= mu * hi6[2]:
                    keep[i] = False                      # fx <= hi implied by fx <= mu fz <= mu fz_max
                if kind == "lo" and c in (0, 1) and lo6[c] <= -mu * hi6[2]:
                    keep[i] = False
                if kind == "fric" and c in (2, 3) and lo6[c - 2] >= 0 and lo6[2] >= 0:
                    keep[i] = False                      # -fx - mu fz <= 0 implied by fx >= 0, fz >= 0
    return rows[keep], rhs[keep], [t for t, k in zip(tags, keep) if k]


def sweep_inverse(M):
    """Inverse of SPD M by the symmetric sweep operator in natural order (what the kernel does)."""
    A = M.copy()
    n = len(A)
    for k in range(n):
        d = 1.0 / A[k, k]
        col = A[:, k].copy()
        A -= np.outer(col, col) * d
        A[:, k] = col * d
        A[k, :] = col * d
        A[k, k] = -d
    return -A


def ipm_model(red, tol=1e-11, maxit=40, prune=True, init_fz_frac=0.1, refine=0, use_sweep=True, verbose=False, chol=False, rd_fac=10.0, abs_tol=None, return_state=False):
    Cb, rb, tags = block_rows(red, prune)
    nb, LB = len(red["blocks"]), red["LB"]
    n, mb = nb * LB, len(rb)
    H, g = red["Hc"], red["g"]
    C = np.zeros((nb * mb, n))
    for j in range(nb):
        C[mb * j:mb * j + mb, LB * j:LB * j + LB] = Cb
    b = np.tile(rb, nb)
    # strictly feasible start inside each block's polytope
    lo6, hi6, comps = red["lo6"], red["hi6"], red["comps"]
    ub = np.zeros(LB)
    for k, c in enumerate(comps):
        ub[k] = 0.5 * (lo6[c] + hi6[c])
    if 2 in comps:
        kz = comps.index(2)
        ub[kz] = lo6[2] + init_fz_frac * (hi6[2] - lo6[2])
        mu_f = -red["Fb"][0][kz]
        for c in (0, 1):
            if c in comps:
                lo_c = max(lo6[c], -mu_f * ub[kz]); hi_c = min(hi6[c], mu_f * ub[kz])
                ub[comps.index(c)] = 0.5 * (lo_c + hi_c)
    u = np.tile(ub, nb)
    s = b - C @ u
    if s.min() <= 0:
        raise RuntimeError("start not strictly feasible: %s" % s.min())
    m = len(b)
    gs = 1.0 + np.abs(g).max()
    lam = np.full(m, 1.0)
    mu0 = float(s @ lam) / m
    lam = mu0 / s * 1.0   # centred start: s*lam = const
    it = 0
    hist = []
    for it in range(1, maxit + 1):
        rd = H @ u + g + C.T @ lam
        rp = C @ u + s - b
        mu = float(s @ lam) / m
        hist.append((mu, np.abs(rd).max(), np.abs(rp).max()))
        if verbose:
            print(it, "mu %.2e rd %.2e rp %.2e" % hist[-1])
        if abs_tol is not None:
            if mu <= abs_tol[0] and np.abs(rd).max() <= abs_tol[1]:
                break
        elif mu <= tol * gs and np.abs(rd).max() <= tol * gs * rd_fac:
            break
        d = lam / s
        M = H + C.T @ (d[:, None] * C)
        if chol:
            import scipy.linalg as sla
            try:
                cf = sla.cho_factor(M, lower=True)
            except Exception:
                hist.append(('cholfail', it, mu))
                break
        else:
            Minv = sweep_inverse(M) if use_sweep else np.linalg.inv(M)

        def solve(rhs):
            if chol:
                return sla.cho_solve(cf, rhs)
            x = Minv @ rhs
            for _ in range(refine):
                x = x + Minv @ (rhs - M @ x)
            return x

        def newton(rc):
            du = solve(-rd - C.T @ (d * rp - rc / s))
            ds = -rp - C @ du
            dl = (-rc - lam * ds) / s
            return du, ds, dl

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0

        du, ds, dl = newton(s * lam)
        a = min(max_step(s, ds), max_step(lam, dl))
        mu_aff = float((s + a * ds) @ (lam + a * dl)) / m
        sigma = (mu_aff / mu) ** 3
        du, ds, dl = newton(s * lam + ds * dl - sigma * mu)
        a = min(1.0, 0.995 * min(max_step(s, ds), max_step(lam, dl)))
        u, s, lam = u + a * du, s + a * ds, lam + a * dl
    if return_state:
        return u, it, hist, s, lam
    return u, it, hist


# ----------------------------------------------------------------------------
# model of the kernel's active-set polish (written the way the CUDA code does it)
# ----------------------------------------------------------------------------

def block_nullspace(A, b, tol=1e-10):
    """Affine set {x : A x = b} of one block by Gauss-Jordan with complete pivoting.

    Returns (p, N, ok): particular solution p (free variables 0), basis N (LB x d) of the null
    space, ok=False if the rows are inconsistent.  A is k x LB with k possibly > LB or rank deficient.
    """
    A = np.array(A, dtype=float, copy=True)
    b = np.array(b, dtype=float, copy=True)
    k, LB = A.shape
    scale = max(1.0, np.abs(A).max()) if k else 1.0
    piv_col = []
    piv_row = []
    used_r = np.zeros(k, bool)
    used_c = np.zeros(LB, bool)
    for _ in range(min(k, LB)):
        best, br, bc = 0.0, -1, -1
        for r in range(k):
            if used_r[r]:
                continue
            for c in range(LB):
                if not used_c[c] and abs(A[r, c]) > best:
                    best, br, bc = abs(A[r, c]), r, c
        if best <= tol * scale:
            break
        used_r[br] = used_c[bc] = True
        piv_row.append(br)
        piv_col.append(bc)
        inv = 1.0 / A[br, bc]
        A[br] *= inv
        b[br] *= inv
        for r in range(k):
            if r != br and A[r, bc] != 0.0:
                f = A[r, bc]
                A[r] -= f * A[br]
                b[r] -= f * b[br]
    ok = True
    for r in range(k):
        if not used_r[r] and abs(b[r]) > 1e-7 * max(1.0, np.abs(b).max()):
            ok = False
    free = [c for c in range(LB) if not used_c[c]]
    p = np.zeros(LB)
    N = np.zeros((LB, len(free)))
    for r, c in zip(piv_row, piv_col):
        p[c] = b[r]
    for jf, c in enumerate(free):
        N[c, jf] = 1.0
        for r, pc in zip(piv_row, piv_col):
            N[pc, jf] = -A[r, c]
    return p, N, ok


def tiny_nnls(A, r, tol=1e-12, maxit=40):
    """Lawson-Hanson: min |A' y - r|, y >= 0, for A (k x LB) with k <= ~18.  Returns y, residual vector."""
    k, LB = A.shape
    y = np.zeros(k)
    passive = np.zeros(k, bool)
    res = r.copy()
    w = A @ res
    scale = max(1.0, np.abs(r).max())
    for _ in range(maxit):
        cand = np.where(~passive, w, -np.inf)
        j = int(np.argmax(cand)) if k else -1
        if k == 0 or cand[j] <= tol * scale * max(1.0, np.abs(A[j]).max()):
            break
        passive[j] = True
        for _inner in range(maxit):
            idx = np.nonzero(passive)[0]
            Ap = A[idx]
            G = Ap @ Ap.T
            G[np.diag_indices_from(G)] += 1e-30
            try:
                z = np.linalg.solve(G, Ap @ r)
            except np.linalg.LinAlgError:
                z = np.linalg.lstsq(Ap.T, r, rcond=None)[0]
            if (z > 0).all():
                y[:] = 0.0
                y[idx] = z
                break
            neg = z <= 0
            alpha = np.min(y[idx][neg] / (y[idx][neg] - z[neg]))
            y[idx] = y[idx] + alpha * (z - y[idx])
            drop = idx[(y[idx] <= 1e-300) | ((z <= 0) & (np.abs(y[idx]) <= 1e-14 * max(1.0, np.abs(y).max())))]
            passive[drop] = False
            y[drop] = 0.0
            if not passive.any():
                break
        res = r - A.T @ y
        w = A @ res
    return y, res


def polish_model(red, u, s, lam, max_rounds=4, verbose=False):
    """Active-set polish from an interior-point iterate.  Returns (u_polished, ok, rounds)."""
    Cb, rb, tags = block_rows(red, True)
    nb, LB = len(red["blocks"]), red["LB"]
    mb = len(rb)
    n = nb * LB
    H, g = red["Hc"], red["g"]
    gs = 1.0 + np.abs(g).max()
    d = lam / s
    hdiag = np.diag(H)
    active = np.zeros((nb, mb), bool)
    for j in range(nb):
        for k in range(mb):
            a = Cb[k]
            eta = (a * a) @ hdiag[LB * j:LB * j + LB] / max((a @ a) ** 2, 1e-300)
            active[j, k] = d[j * mb + k] > eta
    for rnd in range(1, max_rounds + 1):
        P = np.zeros(n)
        Ns = []
        ok = True
        for j in range(nb):
            rows = np.nonzero(active[j])[0]
            p, N, okj = block_nullspace(Cb[rows], rb[rows])
            ok = ok and okj
            P[LB * j:LB * j + LB] = p
            Ns.append(N)
        if not ok:
            return u, False, rnd
        dims = [N.shape[1] for N in Ns]
        nw = sum(dims)
        Nfull = np.zeros((n, nw))
        o = 0
        for j, N in enumerate(Ns):
            Nfull[LB * j:LB * j + LB, o:o + dims[j]] = N
            o += dims[j]
        if nw:
            Hr = Nfull.T @ H @ Nfull
            gr = Nfull.T @ (H @ P + g)
            try:
                w = np.linalg.solve(Hr, -gr)
            except np.linalg.LinAlgError:
                return u, False, rnd
            up = P + Nfull @ w
        else:
            up = P
        changed = False
        # primal check
        for j in range(nb):
            viol = Cb @ up[LB * j:LB * j + LB] - rb
            add = (viol > 1e-9 * (1.0 + np.abs(rb))) & ~active[j]
            if add.any():
                active[j] |= add
                changed = True
        if changed:
            if verbose:
                print("round", rnd, "added violated rows")
            continue
        # dual check
        grad = H @ up + g
        for j in range(nb):
            rows = np.nonzero(active[j])[0]
            rj = -grad[LB * j:LB * j + LB]
            y, res = tiny_nnls(Cb[rows], rj)
            if np.abs(res).max() > 1e-9 * gs:
                wv = Cb[rows] @ res
                drop = rows[(wv < -1e-12 * gs) & (y <= 0)]
                if len(drop) == 0:
                    return u, False, rnd
                active[j, drop] = False
                changed = True
        if not changed:
            return up, True, rnd
        if verbose:
            print("round", rnd, "dropped rows")
    return u, False, max_rounds
