// Active-set polish: from an interior-point iterate to the exact optimum of the QP.
//
// The interior-point iterate identifies the active rows (barrier weight lam/s far above the
// objective curvature along the row).  Those rows become equalities; because every row
// touches one foot-stage block only, the affine set is a product of per-block affine sets:
//     u_b = p_b + N_b w_b        (Gauss-Jordan with complete pivoting, handles dependent rows
//                                 such as the unloaded-foot vertex fx=fy=fz=0)
// and the equality-constrained QP collapses to the unconstrained reduced system
//     (N' Hc N) w = -N' (Hc p + g)
// solved with the same packed Cholesky as the interior-point steps.  The result is accepted
// only with a certificate: primal feasibility of every row and, per block, non-negative
// multipliers found by a tiny Lawson-Hanson NNLS on the active rows (stationarity holds by
// construction in the null space).  Violated rows are added / rows the NNLS residual moves
// away from are released and the polish repeats; failure falls back to more interior-point
// iterations.  Mirrors tools/kernel_model.py::polish_model.
#pragma once
#include "bmpc_kernels.cuh"

namespace bmpc {

// ---- normal equations of the active rows of one block, entirely in registers ---------------
// G = C_A' C_A (LB x LB, PSD) with one right-hand side, reduced by Gauss-Jordan with diagonal
// pivoting.  The pivot is moved to position `step` by predicated swaps, so every register index
// is a compile-time constant (no local memory); perm[] maps positions back to components and the
// first *np positions are the pivots.  A[i][LB] is the transformed right-hand side.
//
// gauss_jordan_diag: A holds the full symmetric matrix and the right-hand side in column LB on entry.
// One elimination step is a function template of the step number (called LB times below): with the step a compile-time
// constant every index into A and perm is one too, and the arrays are scalarised into registers (as a `#pragma unroll` loop
// over the steps the compiler kept the loop and put A and perm into local memory).
template <int LB, int STEP>
BMPC_HD __forceinline__ void gauss_jordan_step(double (&A)[LB][LB + 1], int (&perm)[LB], int& np, bool& done, double scale) {
    int pv = STEP;
    double best = A[STEP][STEP];
#pragma unroll
    for (int c = STEP + 1; c < LB; ++c) {
        const bool gt = A[c][c] > best;
        best = gt ? A[c][c] : best;
        pv = gt ? c : pv;
    }
    const bool take = !done && best > 1e-12 * scale;
    done = done || !take;
#pragma unroll
    for (int c = STEP + 1; c < LB; ++c) {
        const bool sw = take && (pv == c);
#pragma unroll
        for (int col = 0; col <= LB; ++col) {
            const double t = A[STEP][col];
            A[STEP][col] = sw ? A[c][col] : t;
            A[c][col] = sw ? t : A[c][col];
        }
#pragma unroll
        for (int row = 0; row < LB; ++row) {
            const double t = A[row][STEP];
            A[row][STEP] = sw ? A[row][c] : t;
            A[row][c] = sw ? t : A[row][c];
        }
        const int tp = perm[STEP];
        perm[STEP] = sw ? perm[c] : tp;
        perm[c] = sw ? tp : perm[c];
    }
    // (branch-free: a step that is not taken multiplies by 1 and subtracts 0)
    const double inv = take ? 1.0 / A[STEP][STEP] : 1.0;
#pragma unroll
    for (int col = 0; col <= LB; ++col) A[STEP][col] *= inv;
#pragma unroll
    for (int r = 0; r < LB; ++r)
        if (r != STEP) {
            const double f = take ? A[r][STEP] : 0.0;
#pragma unroll
            for (int col = 0; col <= LB; ++col) A[r][col] -= f * A[STEP][col];
        }
    np += take ? 1 : 0;
}
template <int LB>
BMPC_HD __forceinline__ void gauss_jordan_diag(double (&A)[LB][LB + 1], int (&perm)[LB], int& np) {
    static_assert(LB == 5 || LB == 6, "one step call per component below");
    double scale = 1e-300;
#pragma unroll
    for (int a = 0; a < LB; ++a) {
        scale = fmax(scale, A[a][a]);
        perm[a] = a;
    }
    np = 0;
    bool done = false;
    gauss_jordan_step<LB, 0>(A, perm, np, done, scale);
    gauss_jordan_step<LB, 1>(A, perm, np, done, scale);
    gauss_jordan_step<LB, 2>(A, perm, np, done, scale);
    gauss_jordan_step<LB, 3>(A, perm, np, done, scale);
    gauss_jordan_step<LB, 4>(A, perm, np, done, scale);
    if constexpr (LB == 6) gauss_jordan_step<LB, 5>(A, perm, np, done, scale);
}

template <int LB>
BMPC_HD __forceinline__ void normal_reduce(const double* __restrict__ Cb, int mb, unsigned mask,
                                              const double (&rhs)[LB], double (&A)[LB][LB + 1], int (&perm)[LB],
                                              int& np) {
#pragma unroll
    for (int a = 0; a < LB; ++a) {
#pragma unroll
        for (int b = 0; b < LB; ++b) A[a][b] = 0.0;
        A[a][LB] = rhs[a];
    }
#pragma unroll 1
    for (int k = 0; k < mb; ++k)
        if ((mask >> k) & 1u) {
            double cb[LB];
#pragma unroll
            for (int c = 0; c < LB; ++c) cb[c] = Cb[k * LB + c];
#pragma unroll
            for (int a = 0; a < LB; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) A[a][b] += cb[a] * cb[b];
        }
#pragma unroll
    for (int a = 0; a < LB; ++a)
#pragma unroll
        for (int b = 0; b < a; ++b) A[b][a] = A[a][b];
    gauss_jordan_diag<LB>(A, perm, np);
}

// Affine set {x : C_A x = b_A} of the active rows of one block: p[LB] particular solution (free
// components 0), N[c*LB + a] component c of basis vector a (a < *dim_out; unused columns zero).
// p and N are written with run-time component indices, so they should point to shared memory.
// false if the rows are inconsistent.
template <int LB>
BMPC_HD __noinline__ bool block_nullspace(const double* __restrict__ Cb, const double* __restrict__ rb, int mb,
                                             unsigned mask, double* __restrict__ p, double* __restrict__ N,
                                             int* __restrict__ dim_out) {
    double rhs[LB];
#pragma unroll
    for (int c = 0; c < LB; ++c) rhs[c] = 0.0;
    double bmax = 1.0;
#pragma unroll 1
    for (int k = 0; k < mb; ++k)
        if ((mask >> k) & 1u) {
            const double bk = rb[k];
            bmax = fmax(bmax, fabs(bk));
#pragma unroll
            for (int c = 0; c < LB; ++c) rhs[c] += Cb[k * LB + c] * bk;
        }
    double A[LB][LB + 1];
    int perm[LB], np;
    normal_reduce<LB>(Cb, mb, mask, rhs, A, perm, np);
#pragma unroll
    for (int e = 0; e < LB * LB; ++e) N[e] = 0.0;
#pragma unroll
    for (int i = 0; i < LB; ++i) p[perm[i]] = (i < np) ? A[i][LB] : 0.0;
#pragma unroll
    for (int q = 0; q < LB; ++q)
        if (q >= np) {
            const int a = q - np;
#pragma unroll
            for (int i = 0; i < LB; ++i) N[perm[i] * LB + a] = (i < np) ? -A[i][q] : ((i == q) ? 1.0 : 0.0);
        }
    // consistency of the (possibly dependent) active rows
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < mb; ++k)
        if ((mask >> k) & 1u) {
            double acc = -rb[k];
#pragma unroll
            for (int i = 0; i < LB; ++i)
                if (i < np) acc += Cb[k * LB + perm[i]] * A[i][LB];
            if (fabs(acc) > 1e-7 * bmax) ok = false;
        }
    *dim_out = LB - np;
    return ok;
}

// Fast multiplier check.  With lam the interior-point multipliers of the active rows, the
// candidate y = lam + C_A z, G z = r - C_A' lam, is the multiplier vector closest to lam that
// reproduces r = -(gradient) exactly; the interior-point lam sits in the relative interior of the
// optimal dual face, so y >= 0 whenever the active set is right and strictly complementary.
// true: y >= 0 and C_A' y = r (certificate holds).  false: undecided, run the exact NNLS check.
template <int LB>
BMPC_HD __noinline__ bool block_dual_fast(const double* __restrict__ Cb, int mb, unsigned mask,
                                             const double* __restrict__ lam, const double* __restrict__ r, double gs) {
    double rho[LB];
#pragma unroll
    for (int c = 0; c < LB; ++c) rho[c] = r[c];
#pragma unroll 1
    for (int k = 0; k < mb; ++k)
        if ((mask >> k) & 1u) {
            const double lk = lam[k];
#pragma unroll
            for (int c = 0; c < LB; ++c) rho[c] -= Cb[k * LB + c] * lk;
        }
    double A[LB][LB + 1];
    int perm[LB], np;
    normal_reduce<LB>(Cb, mb, mask, rho, A, perm, np);
    double res[LB];
#pragma unroll
    for (int c = 0; c < LB; ++c) res[c] = r[c];
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < mb; ++k)
        if ((mask >> k) & 1u) {
            double y = lam[k];
#pragma unroll
            for (int i = 0; i < LB; ++i)
                if (i < np) y += Cb[k * LB + perm[i]] * A[i][LB];
            if (!(y >= 0.0)) ok = false;
#pragma unroll
            for (int c = 0; c < LB; ++c) res[c] -= Cb[k * LB + c] * y;
        }
    double rmax = 0.0;
#pragma unroll
    for (int c = 0; c < LB; ++c) rmax = fmax(rmax, fabs(res[c]));
    return ok && rmax <= 1e-9 * gs;
}

// Is r (= minus the objective gradient restricted to the block) a non-negative combination of
// the active rows?  Lawson-Hanson NNLS  min |A' y - r|, y >= 0.  Returns true if the residual is
// below tolerance; otherwise *drop receives the active rows the residual direction moves away
// from (to be released), possibly 0 (then the polish gives up).
template <int LB>
BMPC_HD __noinline__ bool block_dual_check(const double* __restrict__ Cb, int mb, unsigned mask, const double* __restrict__ r,
                                 double gs, unsigned* __restrict__ drop) {
    double A[MAXROWS][LB];
    int rows[MAXROWS];
    int k = 0;
#pragma unroll 1
    for (int rr = 0; rr < mb; ++rr)
        if ((mask >> rr) & 1u) {
#pragma unroll 1
            for (int c = 0; c < LB; ++c) A[k][c] = Cb[rr * LB + c];
            rows[k++] = rr;
        }
    double y[MAXROWS];
#pragma unroll 1
    for (int j = 0; j < k; ++j) y[j] = 0.0;
    double res[LB];
    double scale = 1.0;
#pragma unroll 1
    for (int c = 0; c < LB; ++c) res[c] = r[c], scale = fmax(scale, fabs(r[c]));
    unsigned passive = 0u;
#pragma unroll 1
    for (int outer = 0; outer < 40; ++outer) {
        double best = -1e300;
        int bj = -1;
#pragma unroll 1
        for (int j = 0; j < k; ++j) {
            if ((passive >> j) & 1u) continue;
            double w = 0.0;
#pragma unroll 1
            for (int c = 0; c < LB; ++c) w += A[j][c] * res[c];
            if (w > best) best = w, bj = j;
        }
        if (bj < 0) break;
        double amax = 1.0;
#pragma unroll 1
        for (int c = 0; c < LB; ++c) amax = fmax(amax, fabs(A[bj][c]));
        if (best <= 1e-12 * scale * amax) break;
        passive |= 1u << bj;
#pragma unroll 1
        for (int inner = 0; inner < 40; ++inner) {
            int idx[LB + 1], cnt = 0;
#pragma unroll 1
            for (int j = 0; j < k && cnt <= LB; ++j)
                if ((passive >> j) & 1u) idx[cnt++] = j;
            if (cnt > LB) {  // cannot be independent any more: give up on this block
                *drop = 0u;
                return false;
            }
            // normal equations G z = A_P r with partial pivoting
            double G[LB][LB + 1];
#pragma unroll 1
            for (int a = 0; a < cnt; ++a) {
#pragma unroll 1
                for (int b = 0; b < cnt; ++b) {
                    double acc = 0.0;
#pragma unroll 1
                    for (int c = 0; c < LB; ++c) acc += A[idx[a]][c] * A[idx[b]][c];
                    G[a][b] = acc;
                }
                G[a][a] += 1e-30;
                double acc = 0.0;
#pragma unroll 1
                for (int c = 0; c < LB; ++c) acc += A[idx[a]][c] * r[c];
                G[a][cnt] = acc;
            }
#pragma unroll 1
            for (int a = 0; a < cnt; ++a) {
                int pv = a;
#pragma unroll 1
                for (int b = a + 1; b < cnt; ++b)
                    if (fabs(G[b][a]) > fabs(G[pv][a])) pv = b;
                if (pv != a)
#pragma unroll 1
                    for (int c = a; c <= cnt; ++c) {
                        const double t = G[a][c];
                        G[a][c] = G[pv][c];
                        G[pv][c] = t;
                    }
                const double inv = 1.0 / G[a][a];
#pragma unroll 1
                for (int b = a + 1; b < cnt; ++b) {
                    const double f = G[b][a] * inv;
#pragma unroll 1
                    for (int c = a; c <= cnt; ++c) G[b][c] -= f * G[a][c];
                }
            }
            double z[LB];
#pragma unroll 1
            for (int a = cnt - 1; a >= 0; --a) {
                double acc = G[a][cnt];
#pragma unroll 1
                for (int b = a + 1; b < cnt; ++b) acc -= G[a][b] * z[b];
                z[a] = acc / G[a][a];
            }
            bool allpos = true;
#pragma unroll 1
            for (int a = 0; a < cnt; ++a) allpos = allpos && (z[a] > 0.0);
            if (allpos) {
#pragma unroll 1
                for (int a = 0; a < cnt; ++a) y[idx[a]] = z[a];
                break;
            }
            double alpha = 1e300, ymax = 1.0;
#pragma unroll 1
            for (int a = 0; a < cnt; ++a)
                if (z[a] <= 0.0) alpha = fmin(alpha, y[idx[a]] / (y[idx[a]] - z[a]));
            if (!(alpha >= 0.0) || !(alpha < 1e300)) alpha = 0.0;
#pragma unroll 1
            for (int a = 0; a < cnt; ++a) {
                y[idx[a]] += alpha * (z[a] - y[idx[a]]);
                ymax = fmax(ymax, fabs(y[idx[a]]));
            }
#pragma unroll 1
            for (int a = 0; a < cnt; ++a)
                if (y[idx[a]] <= 1e-300 || (z[a] <= 0.0 && fabs(y[idx[a]]) <= 1e-14 * ymax)) {
                    passive &= ~(1u << idx[a]);
                    y[idx[a]] = 0.0;
                }
            if (passive == 0u) break;
        }
#pragma unroll 1
        for (int c = 0; c < LB; ++c) {
            double acc = r[c];
#pragma unroll 1
            for (int j = 0; j < k; ++j) acc -= A[j][c] * y[j];
            res[c] = acc;
        }
    }
    double rmax = 0.0;
#pragma unroll 1
    for (int c = 0; c < LB; ++c) rmax = fmax(rmax, fabs(res[c]));
    if (rmax <= 1e-9 * gs) {
        *drop = 0u;
        return true;
    }
    unsigned d = 0u;
#pragma unroll 1
    for (int j = 0; j < k; ++j) {
        double w = 0.0;
#pragma unroll 1
        for (int c = 0; c < LB; ++c) w += A[j][c] * res[c];
        if (w < -1e-12 * gs && y[j] <= 0.0) d |= 1u << rows[j];
    }
    *drop = d;
    return false;
}

}  // namespace bmpc
