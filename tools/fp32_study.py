"""Measured answer to "why not FP32?" (north_star: tensor cores / lower precision only if shown to hold tolerance; VERDICT
round 1, missing 2).  The interior point of the kernels on the contact-reduced condensed QP (numpy model,
tools/kernel_model.py), with the normal-equations matrix  M = Hc + C' diag(lam/s) C  FACTORED AND SOLVED IN FLOAT32 - plain,
and with 1 / 2 steps of iterative refinement against the FP64 matrix - beside the FP64 factor, on instances drawn from the
bench distribution, stopped at the kernels' loose target.  Reports: Cholesky breakdowns, iterations, the relative error of
the Newton direction against the FP64 direction as the barrier weights grow, and whether the final iterate identifies the
same active set as the FP64 run (what the polish needs).
usage: python tools/fp32_study.py [instances]      (CPU only; writes nothing - the table in profiles/r2_summary.md is its output)"""
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from biped_mpc_py_b200 import synth
from biped_mpc_py_b200.params import MPC, Biped
import kernel_model as km


def ipm(red, mode, refine, tol=1e-7, maxit=40):
    """mode 'f64' | 'f32': precision of the factorisation and of the triangular solves."""
    Cb, rb, _ = km.block_rows(red, True)
    nb, LB = len(red["blocks"]), red["LB"]
    n, mb = nb * LB, len(rb)
    H, g = red["Hc"], red["g"]
    C = np.zeros((nb * mb, n))
    for j in range(nb):
        C[mb * j:mb * j + mb, LB * j:LB * j + LB] = Cb
    b = np.tile(rb, nb)
    lo6, hi6, comps = red["lo6"], red["hi6"], red["comps"]
    ub = np.array([0.5 * (lo6[c] + hi6[c]) for c in comps])
    kz = comps.index(2)
    ub[kz] = lo6[2] + 0.1 * (hi6[2] - lo6[2])
    mu_f = -red["Fb"][0][kz]
    for c in (0, 1):
        ub[comps.index(c)] = 0.5 * (max(lo6[c], -mu_f * ub[kz]) + min(hi6[c], mu_f * ub[kz]))
    u = np.tile(ub, nb)
    s = b - C @ u
    m = len(b)
    gs = 1.0 + np.abs(g).max()
    lam = (float(s.sum()) / m) / s
    info = dict(fail=False, it=0, dir_err=[], dmax=[])
    for it in range(1, maxit + 1):
        rd = H @ u + g + C.T @ lam
        rp = C @ u + s - b
        mu = float(s @ lam) / m
        info["it"] = it
        if mu <= tol * gs and np.abs(rd).max() <= tol * gs * 10.0:
            break
        d = lam / s
        M = H + C.T @ (d[:, None] * C)
        try:
            c64 = sla.cho_factor(M, lower=True)
            cf = sla.cho_factor(M.astype(np.float32), lower=True) if mode == "f32" else c64
        except Exception:
            info["fail"] = True
            break

        def solve(rhs):
            if mode == "f64":
                return sla.cho_solve(cf, rhs)
            x = sla.cho_solve(cf, rhs.astype(np.float32)).astype(np.float64)
            for _ in range(refine):
                x = x + sla.cho_solve(cf, (rhs - M @ x).astype(np.float32)).astype(np.float64)
            return x

        def newton(rc, record=False):
            rhs = -rd - C.T @ (d * rp - rc / s)
            du = solve(rhs)
            if record and mode == "f32":
                ex = sla.cho_solve(c64, rhs)
                info["dir_err"].append(float(np.abs(du - ex).max() / max(1e-300, np.abs(ex).max())))
                info["dmax"].append(float(d.max()))
            ds = -rp - C @ du
            return du, ds, (-rc - lam * ds) / s

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0

        du, ds, dl = newton(s * lam)
        a = min(max_step(s, ds), max_step(lam, dl))
        sigma = (float((s + a * ds) @ (lam + a * dl)) / m / mu) ** 3
        du, ds, dl = newton(s * lam + ds * dl - sigma * mu, record=True)
        a = min(1.0, 0.995 * min(max_step(s, ds), max_step(lam, dl)))
        if not np.isfinite(a) or not np.isfinite(du).all():
            info["fail"] = True
            break
        u, s, lam = u + a * du, s + a * ds, lam + a * dl
    info["converged"] = (not info["fail"]) and info["it"] < maxit
    info["active"] = (lam > s)   # the side of the central path every row ends on
    info["u"] = u
    return info


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    mpc, biped = MPC(), Biped()
    b = synth.make_batch(n, shard_index=77, mpc=mpc, biped=biped)
    variants = [("f64", 0), ("f32", 0), ("f32", 1), ("f32", 2)]
    stats = {v: dict(fail=0, conv=0, its=[], same_active=0, du=[], err_by_d={}) for v in variants}
    for i in range(n):
        red = km.build_reduced(b["x_fb"][i], int(b["phase_k"][i]), b["foot"][i], b["contact"][i], mpc, biped)
        ref = ipm(red, "f64", 0)
        for v in variants:
            r = ref if v == ("f64", 0) else ipm(red, *v)
            st = stats[v]
            st["fail"] += int(r["fail"])
            st["conv"] += int(r["converged"])
            if r["converged"]:
                st["its"].append(r["it"])
                st["same_active"] += int((r["active"] == ref["active"]).all())
                st["du"].append(float(np.abs(r["u"] - ref["u"]).max() / max(1.0, np.abs(ref["u"]).max())))
            for e, dm in zip(r["dir_err"], r["dmax"]):
                st["err_by_d"].setdefault(int(np.floor(np.log10(max(dm, 1.0)))), []).append(e)
    print(f"{n} instances of the bench distribution (85 % walking), interior point stopped at mu <= 1e-7 (1 + |g|):")
    print("| factor / solves | breakdowns | converged | mean iterations | same active set as FP64 | max rel |u - u64| of the iterate |")
    print("|---|---|---|---|---|---|")
    for v in variants:
        st = stats[v]
        name = "FP64" if v[0] == "f64" else f"FP32, {v[1]} refinement step(s)"
        print(f"| {name} | {st['fail']} | {st['conv']} / {n} | {np.mean(st['its']) if st['its'] else float('nan'):.2f} | "
              f"{st['same_active']} / {st['conv']} | {max(st['du']) if st['du'] else float('nan'):.1e} |")
    print("\nrelative error of the FP32 Newton direction (corrector solve) against the FP64 one, by largest barrier weight lam/s of the iteration:")
    print("| max lam/s | " + " | ".join(name for name in ("no refinement", "1 step", "2 steps")) + " |")
    print("|---|---|---|---|")
    decades = sorted(set().union(*[set(stats[v]["err_by_d"]) for v in variants[1:]]))
    for dcd in decades:
        cells = []
        for v in variants[1:]:
            e = stats[v]["err_by_d"].get(dcd)
            cells.append(f"{np.median(e):.1e} (max {np.max(e):.1e})" if e else "-")
        print(f"| 1e{dcd} .. 1e{dcd + 1} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
