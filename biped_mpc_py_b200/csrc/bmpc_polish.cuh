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

// Affine set {x : A x = b} of the active rows of one block.  p[LB] particular solution,
// N[c*LB + a] basis vector a (a < *dim) component c.  false if the rows are inconsistent.
template <int LB>
__device__ __noinline__ bool block_nullspace(const double* __restrict__ Cb, const double* __restrict__ rb, int mb, unsigned mask,
                                double* __restrict__ p, double* __restrict__ N, int* __restrict__ dim_out) {
    double A[MAXROWS][LB + 1];
    int k = 0;
    double scale = 1.0;
#pragma unroll 1
    for (int r = 0; r < mb; ++r)
        if ((mask >> r) & 1u) {
#pragma unroll 1
            for (int c = 0; c < LB; ++c) {
                A[k][c] = Cb[r * LB + c];
                scale = fmax(scale, fabs(A[k][c]));
            }
            A[k][LB] = rb[r];
            ++k;
        }
    unsigned used_r = 0u, used_c = 0u;
    int prow[LB], pcol[LB], np = 0;
#pragma unroll 1
    for (int it = 0; it < LB && it < k; ++it) {
        double best = 0.0;
        int br = -1, bc = -1;
#pragma unroll 1
        for (int r = 0; r < k; ++r) {
            if ((used_r >> r) & 1u) continue;
#pragma unroll 1
            for (int c = 0; c < LB; ++c)
                if (!((used_c >> c) & 1u) && fabs(A[r][c]) > best) best = fabs(A[r][c]), br = r, bc = c;
        }
        if (best <= 1e-10 * scale) break;
        used_r |= 1u << br;
        used_c |= 1u << bc;
        prow[np] = br, pcol[np] = bc, ++np;
        const double inv = 1.0 / A[br][bc];
#pragma unroll 1
        for (int c = 0; c <= LB; ++c) A[br][c] *= inv;
#pragma unroll 1
        for (int r = 0; r < k; ++r) {
            if (r == br) continue;
            const double f = A[r][bc];
            if (f != 0.0)
#pragma unroll 1
                for (int c = 0; c <= LB; ++c) A[r][c] -= f * A[br][c];
        }
    }
    double bmax = 1.0;
#pragma unroll 1
    for (int r = 0; r < k; ++r) bmax = fmax(bmax, fabs(A[r][LB]));
    bool ok = true;
#pragma unroll 1
    for (int r = 0; r < k; ++r)
        if (!((used_r >> r) & 1u) && fabs(A[r][LB]) > 1e-7 * bmax) ok = false;
#pragma unroll 1
    for (int c = 0; c < LB; ++c) p[c] = 0.0;
#pragma unroll 1
    for (int i = 0; i < np; ++i) p[pcol[i]] = A[prow[i]][LB];
    int dim = 0;
#pragma unroll 1
    for (int c = 0; c < LB; ++c) {
        if ((used_c >> c) & 1u) continue;
#pragma unroll 1
        for (int cc = 0; cc < LB; ++cc) N[cc * LB + dim] = 0.0;
        N[c * LB + dim] = 1.0;
#pragma unroll 1
        for (int i = 0; i < np; ++i) N[pcol[i] * LB + dim] = -A[prow[i]][c];
        ++dim;
    }
#pragma unroll 1
    for (int c = dim; c < LB; ++c)  // unused basis columns are zero (the caller pads the reduced system)
#pragma unroll 1
        for (int cc = 0; cc < LB; ++cc) N[cc * LB + c] = 0.0;
    *dim_out = dim;
    return ok;
}

// Is r (= minus the objective gradient restricted to the block) a non-negative combination of
// the active rows?  Lawson-Hanson NNLS  min |A' y - r|, y >= 0.  Returns true if the residual is
// below tolerance; otherwise *drop receives the active rows the residual direction moves away
// from (to be released), possibly 0 (then the polish gives up).
template <int LB>
__device__ __noinline__ bool block_dual_check(const double* __restrict__ Cb, int mb, unsigned mask, const double* __restrict__ r,
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

// y = Hc x + add  for the packed symmetric Hc (one row per thread); result in out[]
__device__ __forceinline__ void symv_packed(const double* __restrict__ Hp, int n, const double* __restrict__ x,
                                            const double* __restrict__ add, double* __restrict__ out) {
    const int tid = threadIdx.x;
    if (tid < n) {
        double acc = add ? add[tid] : 0.0;
        const double* row = Hp + tri(tid, 0);
        for (int j = 0; j <= tid; ++j) acc += row[j] * x[j];
        for (int j = tid + 1; j < n; ++j) acc += Hp[tri(j, tid)] * x[j];
        out[tid] = acc;
    }
}

}  // namespace bmpc
