"""Numpy model of the stage-wise (Riccati) linear algebra backend of the tick kernel.

The contact-reduced QP of tools/kernel_model.py::build_reduced, min 1/2 u'Hc u + g'u, is an LQR problem:
    X_k = A_k X_{k-1} + B_k u_k (+ const),  cost 1/2 sum_k (X_k-xr_k)'Q(X_k-xr_k) + 1/2 u_k'Rb u_k
with A_k = I + dt*E_k (theta += dt*Rinv_k*omega, p += dt*v) and B_k non-zero only in the omega rows (Bw) and
the v rows (Bv).  Instead of the dense n x n Cholesky, systems (Hc + blockdiag(Rt_j - Rb)) du = rhs are solved by
one backward Riccati sweep (factor) and a backward + forward sweep per right-hand side: O(h) instead of O(h^3).
Everything the kernel needs from Hc is provided without forming it: products Hc v + g (rollout + adjoint), diag(Hc).
"""
import numpy as np


class Lqr:
    def __init__(self, red, mpc, biped):
        self.h, self.dt = red["h"], mpc.dt
        self.LB = red["LB"]
        self.blocks = red["blocks"]
        self.nb = len(self.blocks)
        self.Q = np.asarray(mpc.Q, float)[:12]
        R = np.asarray(mpc.R, float)
        comps = red["comps"]
        self.Rb = [np.concatenate([R[3 * l:3 * l + 3], R[6 + 3 * l:9 + 3 * l]])[comps] for (s, l) in self.blocks]
        Pp = red["P"]
        self.Rinv = [None] + [Pp[k] - Pp[k - 1] for k in range(1, self.h)]   # Rinv_0 only enters the free response
        self.W, self.V = red["W"], red["V"]        # (nb,3,LB): omega rows and v rows of B
        self.err = red["e"]                        # free response minus reference, (h,12)
        self.stage_blocks = [[j for j, (s, l) in enumerate(self.blocks) if s == k] for k in range(self.h)]
        self.joseph = False

    # A_k applied from the left / A_k' applied from the left, k >= 1
    def A_mul(self, k, x):
        y = x.copy()
        y[0:3] += self.dt * self.Rinv[k] @ x[6:9]
        y[3:6] += self.dt * x[9:12]
        return y

    def At_mul(self, k, p):
        y = p.copy()
        y[6:9] += self.dt * self.Rinv[k].T @ p[0:3]
        y[9:12] += self.dt * p[3:6]
        return y

    def A_mat(self, k):
        A = np.eye(12)
        A[0:3, 6:9] = self.dt * self.Rinv[k]
        A[3:6, 9:12] = self.dt * np.eye(3)
        return A

    def B_mat(self, k, Bw, Bv):
        js = self.stage_blocks[k]
        B = np.zeros((12, self.LB * len(js)))
        for i, j in enumerate(js):
            B[6:9, self.LB * i:self.LB * i + self.LB] = Bw[j]
            B[9:12, self.LB * i:self.LB * i + self.LB] = Bv[j]
        return B

    # ---- products with the (never formed) condensed Hessian -------------------------------------
    def grad(self, v):
        """Hc v + g  by a forward rollout of the linear response and a backward adjoint sweep."""
        LB, h = self.LB, self.h
        e = np.zeros((h, 12))
        dx = np.zeros(12)
        for k in range(h):
            if k > 0:
                dx = self.A_mul(k, dx)
            for j in self.stage_blocks[k]:
                dx[6:9] += self.W[j] @ v[LB * j:LB * j + LB]
                dx[9:12] += self.V[j] @ v[LB * j:LB * j + LB]
            e[k] = self.err[k] + dx
        out = np.zeros_like(v)
        a = np.zeros(12)
        for k in range(h - 1, -1, -1):
            a = a + self.Q * e[k]
            for j in self.stage_blocks[k]:
                out[LB * j:LB * j + LB] = self.Rb[j] * v[LB * j:LB * j + LB] + self.W[j].T @ a[6:9] + self.V[j].T @ a[9:12]
            if k > 0:
                a = self.At_mul(k, a)
        return out

    def hdiag(self):
        """diag(Hc): Rb + diag(B_j' Pbar_s B_j) with Pbar the cost-to-go of the uncontrolled system."""
        LB, h = self.LB, self.h
        out = np.zeros(self.nb * LB)
        P = np.diag(self.Q)
        for k in range(h - 1, -1, -1):
            for j in self.stage_blocks[k]:
                B = np.vstack([self.W[j], self.V[j]])
                out[LB * j:LB * j + LB] = self.Rb[j] + np.einsum("ac,ab,bc->c", B, P[6:12, 6:12], B)
            if k > 0:
                A = self.A_mat(k)
                P = np.diag(self.Q) + A.T @ P @ A
        return out

    # ---- factor / solve --------------------------------------------------------------------------
    def factor(self, Rt, Bw=None, Bv=None):
        """Backward Riccati sweep for M = Hc - blockdiag(Rb) + blockdiag(Rt) (Rt: list of LB x LB SPD blocks).
        Bw/Bv default to the problem's own input maps (the polish passes N-transformed ones)."""
        Bw = self.W if Bw is None else Bw
        Bv = self.V if Bv is None else Bv
        LB, h = self.LB, self.h
        P = np.diag(self.Q)
        self.K, self.Ginv, self.fBw, self.fBv = [None] * h, [None] * h, Bw, Bv
        for k in range(h - 1, -1, -1):
            js = self.stage_blocks[k]
            if js:
                B = self.B_mat(k, Bw, Bv)
                PB = P @ B
                G = B.T @ PB
                for i, j in enumerate(js):
                    G[LB * i:LB * i + LB, LB * i:LB * i + LB] += Rt[j]
                try:
                    Lc = np.linalg.cholesky(0.5 * (G + G.T))
                except np.linalg.LinAlgError:
                    return False
                Li = np.linalg.solve(Lc, np.eye(len(G)))   # triangular solve on the identity
                Ginv = Li.T @ Li
                F = PB.T @ self.A_mat(k) if k > 0 else np.zeros((len(js) * LB, 12))
                self.Ginv[k], self.K[k] = Ginv, Ginv @ F
                if k > 0:
                    A = self.A_mat(k)
                    if self.joseph:
                        # Joseph form: a sum of positive semi-definite terms (no cancellation)
                        Rblk = np.zeros_like(G)
                        for i, j in enumerate(js):
                            Rblk[LB * i:LB * i + LB, LB * i:LB * i + LB] = Rt[j]
                        Acl = A - B @ self.K[k]
                        P = np.diag(self.Q) + Acl.T @ P @ Acl + self.K[k].T @ Rblk @ self.K[k]
                    else:
                        P = np.diag(self.Q) + A.T @ P @ A - F.T @ self.K[k]
            elif k > 0:
                A = self.A_mat(k)
                P = np.diag(self.Q) + A.T @ P @ A
            P = 0.5 * (P + P.T)
        return True

    def solve(self, rhs):
        """du with M du = rhs, using the stored factor."""
        LB, h = self.LB, self.h
        d = [None] * h
        p = np.zeros(12)
        for k in range(h - 1, -1, -1):
            js = self.stage_blocks[k]
            if js:
                B = self.B_mat(k, self.fBw, self.fBv)
                t = -np.concatenate([rhs[LB * j:LB * j + LB] for j in js]) + B.T @ p
                d[k] = self.Ginv[k] @ t
                if k > 0:
                    p = self.At_mul(k, p) - self.K[k].T @ t
            elif k > 0:
                p = self.At_mul(k, p)
        du = np.zeros_like(rhs)
        dx = np.zeros(12)
        for k in range(h):
            if k > 0:
                dx_in = dx
                dx = self.A_mul(k, dx)
            js = self.stage_blocks[k]
            if js:
                uk = -d[k] - (self.K[k] @ dx_in if k > 0 else 0.0)
                B = self.B_mat(k, self.fBw, self.fBv)
                dx = dx + B @ uk
                for i, j in enumerate(js):
                    du[LB * j:LB * j + LB] = uk[LB * i:LB * i + LB]
        return du


if __name__ == "__main__":
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import reference_mpc as rm
    from tools import kernel_model as km
    from biped_mpc_py_b200 import synth, MPC
    rng = np.random.default_rng(0)
    for h in (10, 30):
        mpc, biped = rm.MPCParams(h=h), rm.BipedParams()
        b = synth.make_batch(8, shard_index=21, mpc=MPC(h=h), extend=True, walking_prob=0.6)
        for i in range(8):
            red = km.build_reduced(b["x_fb"][i], int(b["phase_k"][i]), b["foot"][i], b["contact"][i], mpc, biped, extend=True)
            lq = Lqr(red, mpc, biped)
            n, LB = len(red["g"]), red["LB"]
            v = rng.normal(size=n) * 50
            e1 = np.abs(lq.grad(v) - (red["Hc"] @ v + red["g"])).max() / np.abs(red["Hc"] @ v + red["g"]).max()
            e2 = np.abs(lq.hdiag() - np.diag(red["Hc"])).max() / np.diag(red["Hc"]).max()
            # barrier-like block-diagonal terms spanning many decades
            Rt, M = [], red["Hc"].copy()
            for j in range(lq.nb):
                Cj = rng.normal(size=(11, LB))
                D = 10.0 ** rng.uniform(-6, 8, 11)
                blk = np.diag(lq.Rb[j]) + Cj.T @ (D[:, None] * Cj)
                Rt.append(blk)
                M[LB * j:LB * j + LB, LB * j:LB * j + LB] += blk - np.diag(lq.Rb[j])
            assert lq.factor(Rt)
            rhs = rng.normal(size=n) * 1e3
            x_ref = np.linalg.solve(M, rhs)
            x = lq.solve(rhs)
            e3 = np.abs(x - x_ref).max() / np.abs(x_ref).max()
            r3 = np.abs(M @ x - rhs).max() / np.abs(rhs).max()
            print(f"h={h} inst {i} gait {b['gait'][i]} n={n}: grad {e1:.1e} hdiag {e2:.1e} solve err {e3:.1e} resid {r3:.1e} cond(M) {np.linalg.cond(M):.1e}")
