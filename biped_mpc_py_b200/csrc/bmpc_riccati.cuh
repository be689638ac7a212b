// Stage-wise (Riccati) linear-algebra backend of the tick kernel.
//
// The contact-reduced QP is an LQR problem in disguise:  X_k = A_k X_{k-1} + B_k u_k (+ const) with
// A_k = I + dt E_k (theta += dt Rinv_k omega, p += dt v; MPC.py:165-171, 183) and B_k non-zero only in the
// omega rows (Bw = dt Iw^-1 [skew(r) | I]) and the v rows (Bv = dt/m [I | 0]) (MPC.py:174-184), cost
// 1/2 sum_k (X_k - xref_k)' Q (X_k - xref_k) + 1/2 u_k' R u_k (MPC.py:277-286).  Every system the interior point
// and the polish solve has the form (Hc - blockdiag(Rb) + blockdiag(Rt_j)) du = rhs with per-block Rt_j, i.e.
// the same LQR with modified input weights, so ONE backward Riccati sweep factors it and a backward + forward
// sweep solves each right-hand side: O(h) work and storage instead of the O(h^3) / O(h^2) of the dense
// condensed factor.  Products Hc v + g and diag(Hc) are evaluated without forming Hc (rollout + adjoint,
// uncontrolled cost-to-go).  Model and derivation: tools/riccati_model.py (checked against the dense form).
//
// Every stage has 0, 1 or 2 stance feet; the per-stage code is instantiated for CNT = 1 and 2 so that every index
// computation divides by a compile-time constant.  Phases are separated by the group barrier (__syncwarp for a
// one-warp group).
#pragma once
#include "bmpc_kernels.cuh"

namespace bmpc {

// row of entry e of the packed lower triangle of a 12 x 12 matrix
__constant__ unsigned char TRI12_ROW[78] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11};

template <int NT>
__device__ __forceinline__ void rsync() {
    if constexpr (NT == 32) __syncwarp();
    else __syncthreads();
}

// Shared-memory layout of one robot's Riccati data: compile-time offsets from ONE base pointer (a struct of
// pointers passed by reference lives in local memory and costs an LDL per access).
template <int LB, int HZ, int SMAX>
struct Ric {
    static constexpr int NU = 2 * LB;            // inputs per stage: at most two stance feet
    static constexpr int NAB = LB * (LB + 1) / 2;
    static constexpr int HZc = HZ;
    static constexpr int oK = 0;                             // [S][LB][12]  feedback rows of the block's inputs (K = Ginv F)
    static constexpr int oGi = oK + SMAX * LB * 12;          // [S][LB][NU]  the block's rows of inv(G_k)
    static constexpr int oW0 = oGi + SMAX * LB * NU;         // [S][3][LB]   omega rows of B for the block (problem data)
    static constexpr int oBw = oW0 + SMAX * 3 * LB;          // [S][3][LB]   effective omega rows (= W0, or W0 N in the polish)
    static constexpr int oBv = oBw + SMAX * 3 * LB;          // [S][3][LB]   effective v rows (= dt/m selector, or that times N)
    static constexpr int oRt = oBv + SMAX * 3 * LB;          // [S][NAB]     lower triangle of the block's input weight
    static constexpr int orinv = oRt + SMAX * NAB;           // [HZ][9]
    static constexpr int oerr = orinv + HZ * 9;              // [HZ][12]     free response minus reference
    static constexpr int oet = oerr + HZ * 12;               // [HZ][12]     trajectory scratch of the gradient
    static constexpr int oP = oet + HZ * 12;                 // [12][12]
    static constexpr int oPA = oP + 144;                     // [12][12]
    static constexpr int oPB = oPA + 144;                    // [12][NU]
    static constexpr int oF = oPB + 12 * NU;                 // [NU][12]
    static constexpr int oG = oF + NU * 12;                  // [NU][NU]    (lower triangle used by the sweep)
    static constexpr int ovec = oG + NU * NU;                // [64]: p/dx double buffers (2 x 12), t (NU)
    static constexpr int total = ovec + 64;
    double* b;           // base
    const int* sfirst;   // [HZ] first block of the stage
    const int* scnt;     // [HZ] blocks in the stage (0, 1, 2)
    const int* blk_foot;
    double dt, vm;       // vm = dt / mass
    __device__ __forceinline__ double* K() const { return b + oK; }
    __device__ __forceinline__ double* Gi() const { return b + oGi; }
    __device__ __forceinline__ double* W0() const { return b + oW0; }
    __device__ __forceinline__ double* Bw() const { return b + oBw; }
    __device__ __forceinline__ double* Bv() const { return b + oBv; }
    __device__ __forceinline__ double* Rt() const { return b + oRt; }
    __device__ __forceinline__ double* rinv() const { return b + orinv; }
    __device__ __forceinline__ double* err() const { return b + oerr; }
    __device__ __forceinline__ double* et() const { return b + oet; }
    __device__ __forceinline__ double* P() const { return b + oP; }
    __device__ __forceinline__ double* PA() const { return b + oPA; }
    __device__ __forceinline__ double* PB() const { return b + oPB; }
    __device__ __forceinline__ double* F() const { return b + oF; }
    __device__ __forceinline__ double* G() const { return b + oG; }
    __device__ __forceinline__ double* vec() const { return b + ovec; }
};

// block-diagonal input weight Rt_j: interior point  Rb + Cb' diag(d_j) Cb,  polish  N_j' Rb N_j (+ I on the padding)
struct RtSpec {
    int polish;
    const double* Cb;   // [mb][LB]
    const double* dd;   // [S*mb] barrier weights lam/s
    int mb;
    const double* Nn;   // [S][LB*LB] null-space blocks (component c, basis vector a at c*LB + a)
    const int* bdim;    // [S]
};

template <int LB>
__device__ __forceinline__ double block_R(const DevParams& p, int foot, int c) {  // R entry of free component c (MPC.py:28)
    const int ca = p.comps[c];
    return p.R[(ca < 3) ? (3 * foot + ca) : (6 + 3 * foot + ca - 3)];
}

// effective input maps: interior point (Bw = W0, Bv = dt/m selector) or polish (both times N_j)
template <int LB, int NT, class R>
__device__ __noinline__ void ric_set_maps(const DevParams& p, const R r, int S, const double* Nn /* null: interior point */) {
    const int tid = threadIdx.x & (NT - 1);
    for (int e = tid; e < S * 3 * LB; e += NT) {
        const int j = e / (3 * LB), x = (e - j * 3 * LB) / LB, a = e % LB;
        double bw, bv;
        if (Nn == nullptr) {
            bw = r.W0()[e];
            bv = (p.comps[a] == x) ? r.vm : 0.0;
        } else {
            const double* N = Nn + j * LB * LB;
            bw = 0.0, bv = 0.0;
#pragma unroll
            for (int c = 0; c < LB; ++c) {
                bw += r.W0()[(j * 3 + x) * LB + c] * N[c * LB + a];
                if (p.comps[c] == x) bv += r.vm * N[c * LB + a];
            }
        }
        r.Bw()[e] = bw;
        r.Bv()[e] = bv;
    }
    rsync<NT>();
}

// PA = P A_k, then P <- Q + A_k' PA - F'K (the F'K term only when NUK > 0).  A_k = I + dt E_k.
template <int LB, int NT, int NUK, class R>
__device__ __forceinline__ void ric_step_P(const DevParams& p, const R r, int k, int j0) {
    const int tid = threadIdx.x & (NT - 1);
    const double* ri = r.rinv() + 9 * k;
    const double dt = r.dt;
#pragma unroll 1
    for (int e = tid; e < 144; e += NT) {
        const int a = e / 12, b = e - 12 * a;
        double v = r.P()[e];
        if (b >= 6 && b < 9) v += dt * (r.P()[a * 12 + 0] * ri[b - 6] + r.P()[a * 12 + 1] * ri[3 + b - 6] + r.P()[a * 12 + 2] * ri[6 + b - 6]);
        if (b >= 9) v += dt * r.P()[a * 12 + 3 + b - 9];
        r.PA()[e] = v;
    }
    rsync<NT>();
    // lower triangle, mirrored: keeps P exactly symmetric
    const double* Kb = r.K() + j0 * LB * 12;  // the stage's K rows are contiguous: [NUK][12]
#pragma unroll 1
    for (int e = tid; e < 78; e += NT) {
        const int a = TRI12_ROW[e], b = e - a * (a + 1) / 2;  // e = a(a+1)/2 + b, b <= a (table instead of a float square root)
        double v = r.PA()[a * 12 + b];
        if (a >= 6 && a < 9) v += dt * (ri[a - 6] * r.PA()[b] + ri[3 + a - 6] * r.PA()[12 + b] + ri[6 + a - 6] * r.PA()[24 + b]);
        if (a >= 9) v += dt * r.PA()[(3 + a - 9) * 12 + b];
        if (a == b) v += p.Q[a];
#pragma unroll
        for (int c = 0; c < NUK; ++c) v -= r.F()[c * 12 + a] * Kb[c * 12 + b];
        r.P()[a * 12 + b] = v;
        r.P()[b * 12 + a] = v;
    }
    rsync<NT>();
}

// one stage of the backward sweep with CNT stance feet (nu = CNT * LB inputs)
template <int LB, int NT, int CNT, class R>
__device__ __forceinline__ bool ric_factor_stage(const DevParams& p, const R r, int k, int j0) {
    constexpr int NU = 2 * LB, nu = CNT * LB, NAB = LB * (LB + 1) / 2;
    const int tid = threadIdx.x & (NT - 1);
    // PB = P B_k
#pragma unroll 1
    for (int e = tid; e < 12 * nu; e += NT) {
        const int a = e / nu, c = e - a * nu, j = j0 + c / LB, cc = c % LB;
        const double* bw = r.Bw() + j * 3 * LB + cc;
        const double* bv = r.Bv() + j * 3 * LB + cc;
        const double* Pa = r.P() + a * 12;
        r.PB()[a * NU + c] = Pa[6] * bw[0] + Pa[7] * bw[LB] + Pa[8] * bw[2 * LB] + Pa[9] * bv[0] + Pa[10] * bv[LB] + Pa[11] * bv[2 * LB];
    }
    rsync<NT>();
    // G = Rt + B_k' PB (lower triangle) ;  F = PB' A_k
    constexpr int ntri = nu * (nu + 1) / 2;
#pragma unroll 1
    for (int e = tid; e < ntri; e += NT) {
        int rr = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        if (rr * (rr + 1) / 2 > e) --rr;
        if ((rr + 1) * (rr + 2) / 2 <= e) ++rr;
        const int c = e - rr * (rr + 1) / 2;  // c <= rr
        const int jr = j0 + rr / LB, ra = rr % LB, jc = j0 + c / LB, ca = c % LB;
        const double* bw = r.Bw() + jr * 3 * LB + ra;
        const double* bv = r.Bv() + jr * 3 * LB + ra;
        double v = bw[0] * r.PB()[6 * NU + c] + bw[LB] * r.PB()[7 * NU + c] + bw[2 * LB] * r.PB()[8 * NU + c] +
                   bv[0] * r.PB()[9 * NU + c] + bv[LB] * r.PB()[10 * NU + c] + bv[2 * LB] * r.PB()[11 * NU + c];
        if (jr == jc) v += r.Rt()[jr * NAB + ra * (ra + 1) / 2 + ca];
        r.G()[rr * NU + c] = v;
    }
    if (k > 0) {
        const double* ri = r.rinv() + 9 * k;
#pragma unroll 1
        for (int e = tid; e < nu * 12; e += NT) {
            const int c = e / 12, a = e - 12 * c;
            double v = r.PB()[a * NU + c];
            if (a >= 6 && a < 9) v += r.dt * (r.PB()[0 * NU + c] * ri[a - 6] + r.PB()[1 * NU + c] * ri[3 + a - 6] + r.PB()[2 * NU + c] * ri[6 + a - 6]);
            if (a >= 9) v += r.dt * r.PB()[(3 + a - 9) * NU + c];
            r.F()[c * 12 + a] = v;
        }
    }
    rsync<NT>();
    // symmetric sweep of every pivot on the lower triangle: G <- -inv(G)  (SPD: no pivoting).  <= 55 entries and 2 nu
    // dependent steps: run by the FIRST WARP alone with warp barriers (a CTA barrier per step costs more than the
    // three idle warps could contribute)
    bool ok = true;
    if (NT == 32 || (threadIdx.x & (NT - 1)) < 32) {
        const int lane = threadIdx.x & 31;
        constexpr int NR = (ntri + 31) / 32;
        // this lane's entries (i >= j) of the lower triangle: decoded once, the same for every pivot
        int ei[NR], ej[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const int e = lane + q * 32;
            int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
            if (i * (i + 1) / 2 > e) --i;
            if ((i + 1) * (i + 2) / 2 <= e) ++i;
            ei[q] = (e < ntri) ? i : -1;
            ej[q] = e - i * (i + 1) / 2;
        }
        double* Gs = r.G();
#pragma unroll 1
        for (int pv = 0; pv < nu; ++pv) {
            const double gpp = Gs[pv * NU + pv];
            ok = ok && (gpp > 0.0) && (gpp < 1e300);
            const double d = 1.0 / gpp;
            double nv[NR];
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                const int i = ei[q], j = ej[q];
                if (i >= 0) {
                    const double gip = (i >= pv) ? Gs[i * NU + pv] : Gs[pv * NU + i];
                    const double gpj = (j >= pv) ? Gs[j * NU + pv] : Gs[pv * NU + j];
                    const double gij = Gs[i * NU + j];
                    nv[q] = (i == pv && j == pv) ? -d : ((i == pv || j == pv) ? ((i == pv) ? gpj : gip) * d : gij - gip * gpj * d);
                }
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < NR; ++q)
                if (ei[q] >= 0) Gs[ei[q] * NU + ej[q]] = nv[q];
            __syncwarp();
        }
    }
    if constexpr (NT == 32) __syncwarp();
    else ok = __syncthreads_and(ok ? 1 : 0) != 0;
    if (!ok) return false;  // uniform
    // rows of inv(G) = -G (full rows, from the lower triangle); K = inv(G) F
    double* Gib = r.Gi() + j0 * LB * NU;  // the stage's rows are contiguous: [nu][NU]
#pragma unroll 1
    for (int e = tid; e < nu * nu; e += NT) {
        const int rr = e / nu, c = e - rr * nu;
        Gib[rr * NU + c] = -((rr >= c) ? r.G()[rr * NU + c] : r.G()[c * NU + rr]);
    }
    rsync<NT>();
    if (k > 0) {
        double* Kb = r.K() + j0 * LB * 12;
#pragma unroll 1
        for (int e = tid; e < nu * 12; e += NT) {
            const int rr = e / 12, a = e - 12 * rr;
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < nu; ++c) v += Gib[rr * NU + c] * r.F()[c * 12 + a];
            Kb[rr * 12 + a] = v;
        }
        rsync<NT>();
        ric_step_P<LB, NT, nu>(p, r, k, j0);
    }
    return true;
}

// Backward Riccati sweep: stores K and inv(G) rows per block.  Uniform return value (false: a pivot was not positive).
template <int LB, int NT, class R>
__device__ __noinline__ bool ric_factor(const DevParams& p, const R r, const RtSpec& rt, int S) {
    constexpr int NAB = LB * (LB + 1) / 2;
    const int tid = threadIdx.x & (NT - 1);
    for (int e = tid; e < 144; e += NT) r.P()[e] = (e / 12 == e % 12) ? p.Q[e / 12] : 0.0;
    // block input weights (lower triangles), once per factorisation
#pragma unroll 1
    for (int e = tid; e < S * NAB; e += NT) {
        const int j = e / NAB, ab = e - j * NAB;
        int a = 0, b = ab;  // ab = a(a+1)/2 + b
        while (b > a) ++a, b -= a;
        const int foot = r.blk_foot[j];
        double acc = 0.0;
        if (!rt.polish) {
            if (a == b) acc = block_R<LB>(p, foot, a);
            const double* dd = rt.dd + j * rt.mb;
#pragma unroll 1
            for (int k = 0; k < rt.mb; ++k) acc += rt.Cb[k * LB + a] * rt.Cb[k * LB + b] * dd[k];
        } else {
            const double* N = rt.Nn + j * LB * LB;
#pragma unroll
            for (int c = 0; c < LB; ++c) acc += N[c * LB + a] * block_R<LB>(p, foot, c) * N[c * LB + b];
            if (a == b && a >= rt.bdim[j]) acc += 1.0;
        }
        r.Rt()[e] = acc;
    }
    rsync<NT>();
    for (int k = R::HZc - 1; k >= 0; --k) {
        const int cnt = r.scnt[k], j0 = r.sfirst[k];
        bool ok = true;
        if (cnt == 1) ok = ric_factor_stage<LB, NT, 1>(p, r, k, j0);
        else if (cnt == 2) ok = ric_factor_stage<LB, NT, 2>(p, r, k, j0);
        else if (k > 0) ric_step_P<LB, NT, 0>(p, r, k, j0);
        if (!ok) return false;
    }
    return true;
}

template <int LB, int NT, int CNT, class R>
__device__ __forceinline__ void ric_solve_back_stage(const R r, double* __restrict__ x, int k, int j0, const double* pv, double* pn,
                                                     double* tv) {
    constexpr int NU = 2 * LB, nu = CNT * LB;
    const int tid = threadIdx.x & (NT - 1);
    const double* Gib = r.Gi() + j0 * LB * NU;
    const double* Kb = r.K() + j0 * LB * 12;
    if constexpr (CNT > 0) {
        for (int c = tid; c < nu; c += NT) {
            const int j = j0 + c / LB, cc = c % LB;
            const double* bw = r.Bw() + j * 3 * LB + cc;
            const double* bv = r.Bv() + j * 3 * LB + cc;
            tv[c] = -x[j0 * LB + c] + bw[0] * pv[6] + bw[LB] * pv[7] + bw[2 * LB] * pv[8] + bv[0] * pv[9] + bv[LB] * pv[10] +
                    bv[2 * LB] * pv[11];
        }
        rsync<NT>();
    }
    // lanes [0, nu): d = inv(G) t (stored in x);  lanes [16, 28) (or the same lanes when the group is small): p <- A' p - K' t
    constexpr int off = (NT >= 32) ? 16 : 0;
    for (int c = tid; c < nu; c += NT) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < nu; ++q) v += Gib[c * NU + q] * tv[q];
        x[j0 * LB + c] = v;
    }
    if (k > 0) {
        const double* ri = r.rinv() + 9 * k;
        for (int a = tid - off; a < 12; a += NT) {
            if (a < 0) continue;
            double v = pv[a];
            if (a >= 6 && a < 9) v += r.dt * (ri[a - 6] * pv[0] + ri[3 + a - 6] * pv[1] + ri[6 + a - 6] * pv[2]);
            if (a >= 9) v += r.dt * pv[3 + a - 9];
#pragma unroll
            for (int c = 0; c < nu; ++c) v -= Kb[c * 12 + a] * tv[c];
            pn[a] = v;
        }
    }
    rsync<NT>();
}

template <int LB, int NT, int CNT, class R>
__device__ __forceinline__ void ric_solve_fwd_stage(const R r, double* __restrict__ x, int k, int j0, const double* dx, double* dn) {
    constexpr int nu = CNT * LB;
    const int tid = threadIdx.x & (NT - 1);
    const double* Kb = r.K() + j0 * LB * 12;
    if constexpr (CNT > 0) {
        for (int c = tid; c < nu; c += NT) {
            double v = -x[j0 * LB + c];
            if (k > 0) {
#pragma unroll
                for (int a = 0; a < 12; ++a) v -= Kb[c * 12 + a] * dx[a];
            }
            x[j0 * LB + c] = v;
        }
        rsync<NT>();
    }
    const double* ri = r.rinv() + 9 * k;
    for (int a = tid; a < 12; a += NT) {
        double v = dx[a];
        if (k > 0) {
            if (a < 3) v += r.dt * (ri[3 * a] * dx[6] + ri[3 * a + 1] * dx[7] + ri[3 * a + 2] * dx[8]);
            else if (a < 6) v += r.dt * dx[9 + a - 3];
        }
        if (a >= 6) {
            const double* bm = ((a < 9) ? r.Bw() : r.Bv()) + (j0 * 3 + ((a < 9) ? a - 6 : a - 9)) * LB;
#pragma unroll
            for (int c = 0; c < nu; ++c) v += bm[(c / LB) * 3 * LB + c % LB] * x[j0 * LB + c];
        }
        dn[a] = v;
    }
    rsync<NT>();
}

// x <- inv(M) x with the factor of ric_factor (x: LB*S entries in shared memory, block order).  Every phase has at
// most 28 active lanes and the phases are strictly dependent, so the whole solve runs on the FIRST WARP with warp
// barriers; the other warps of a larger group wait at one CTA barrier.
template <int LB, int NT, class R>
__device__ __noinline__ void ric_solve(const R r, double* __restrict__ x) {
    if constexpr (NT > 32) __syncthreads();
    if (NT == 32 || (threadIdx.x & (NT - 1)) < 32) {
        constexpr int WT = 32;
        const int tid = threadIdx.x & 31;
        double* pv = r.vec();        // [12] adjoint
        double* pn = r.vec() + 12;   // [12] next
        double* tv = r.vec() + 36;   // [NU]
        for (int a = tid; a < 12; a += WT) pv[a] = 0.0;
        __syncwarp();
        // backward: t_k = -rhs_k + B_k' p,  d_k = inv(G_k) t_k (stored in x),  p <- A_k' p - K_k' t_k
        for (int k = R::HZc - 1; k >= 0; --k) {
            const int cnt = r.scnt[k], j0 = r.sfirst[k];
            if (cnt == 1) ric_solve_back_stage<LB, WT, 1>(r, x, k, j0, pv, pn, tv);
            else if (cnt == 2) ric_solve_back_stage<LB, WT, 2>(r, x, k, j0, pv, pn, tv);
            else ric_solve_back_stage<LB, WT, 0>(r, x, k, j0, pv, pn, tv);
            if (k > 0) {
                double* t = pv;
                pv = pn;
                pn = t;
            }
        }
        // forward: du_k = -K_k dx_{k-1} - d_k,  dx_k = A_k dx_{k-1} + B_k du_k
        double* dx = r.vec();
        double* dn = r.vec() + 12;
        for (int a = tid; a < 12; a += WT) dx[a] = 0.0;
        __syncwarp();
        for (int k = 0; k < R::HZc; ++k) {
            const int cnt = r.scnt[k], j0 = r.sfirst[k];
            if (cnt == 1) ric_solve_fwd_stage<LB, WT, 1>(r, x, k, j0, dx, dn);
            else if (cnt == 2) ric_solve_fwd_stage<LB, WT, 2>(r, x, k, j0, dx, dn);
            else ric_solve_fwd_stage<LB, WT, 0>(r, x, k, j0, dx, dn);
            double* t = dx;
            dx = dn;
            dn = t;
        }
    }
    if constexpr (NT > 32) __syncthreads();
}

// out = Hc v + g  (rollout of the linear response on top of the free response, then the adjoint sweep)
template <int LB, int NT, class R>
__device__ __noinline__ void ric_grad(const DevParams& p, const R r, const double* __restrict__ v, double* __restrict__ out) {
    const int tid = threadIdx.x & (NT - 1);
    double* dx = r.vec();
    double* dn = r.vec() + 12;
    for (int a = tid; a < 12; a += NT) dx[a] = 0.0;
    rsync<NT>();
    for (int k = 0; k < R::HZc; ++k) {
        const int cnt = r.scnt[k], j0 = r.sfirst[k], nu = cnt * LB;
        const double* ri = r.rinv() + 9 * k;
        for (int a = tid; a < 12; a += NT) {
            double s = dx[a];
            if (k > 0) {
                if (a < 3) s += r.dt * (ri[3 * a] * dx[6] + ri[3 * a + 1] * dx[7] + ri[3 * a + 2] * dx[8]);
                else if (a < 6) s += r.dt * dx[9 + a - 3];
            }
            if (a >= 6) {
#pragma unroll 1
                for (int c = 0; c < nu; ++c) {
                    const int j = j0 + c / LB, cc = c % LB;
                    const double b = (a < 9) ? r.W0()[(j * 3 + a - 6) * LB + cc] : ((p.comps[cc] == a - 9) ? r.vm : 0.0);
                    s += b * v[j0 * LB + c];
                }
            }
            dn[a] = s;
            r.et()[12 * k + a] = r.err()[12 * k + a] + s;
        }
        rsync<NT>();
        double* t = dx;
        dx = dn;
        dn = t;
    }
    double* av = r.vec();
    double* an = r.vec() + 12;
    for (int a = tid; a < 12; a += NT) av[a] = 0.0;
    rsync<NT>();
    for (int k = R::HZc - 1; k >= 0; --k) {
        const int cnt = r.scnt[k], j0 = r.sfirst[k], nu = cnt * LB;
        for (int a = tid; a < 12; a += NT) av[a] += p.Q[a] * r.et()[12 * k + a];
        rsync<NT>();
        for (int c = tid; c < nu; c += NT) {
            const int j = j0 + c / LB, cc = c % LB;
            double s = block_R<LB>(p, r.blk_foot[j], cc) * v[j0 * LB + c] + r.W0()[(j * 3 + 0) * LB + cc] * av[6] +
                       r.W0()[(j * 3 + 1) * LB + cc] * av[7] + r.W0()[(j * 3 + 2) * LB + cc] * av[8];
            if (p.comps[cc] < 3) s += r.vm * av[9 + p.comps[cc]];
            out[j0 * LB + c] = s;
        }
        if (k > 0) {
            const double* ri = r.rinv() + 9 * k;
            for (int a = tid; a < 12; a += NT) {
                double s = av[a];
                if (a >= 6 && a < 9) s += r.dt * (ri[a - 6] * av[0] + ri[3 + a - 6] * av[1] + ri[6 + a - 6] * av[2]);
                if (a >= 9) s += r.dt * av[3 + a - 9];
                an[a] = s;
            }
            rsync<NT>();
            double* t = av;
            av = an;
            an = t;
        } else {
            rsync<NT>();
        }
    }
}

// out = diag(Hc): Rb + diag(B_j' Pbar B_j), Pbar = cost-to-go of the uncontrolled system
template <int LB, int NT, class R>
__device__ __noinline__ void ric_hdiag(const DevParams& p, const R r, double* __restrict__ out) {
    const int tid = threadIdx.x & (NT - 1);
    for (int e = tid; e < 144; e += NT) r.P()[e] = (e / 12 == e % 12) ? p.Q[e / 12] : 0.0;
    rsync<NT>();
    for (int k = R::HZc - 1; k >= 0; --k) {
        const int cnt = r.scnt[k], j0 = r.sfirst[k], nu = cnt * LB;
        for (int c = tid; c < nu; c += NT) {
            const int j = j0 + c / LB, cc = c % LB;
            double b[6];
#pragma unroll
            for (int x = 0; x < 3; ++x) b[x] = r.W0()[(j * 3 + x) * LB + cc], b[3 + x] = (p.comps[cc] == x) ? r.vm : 0.0;
            double s = block_R<LB>(p, r.blk_foot[j], cc);
#pragma unroll
            for (int x = 0; x < 6; ++x)
#pragma unroll
                for (int y = 0; y < 6; ++y) s += b[x] * r.P()[(6 + x) * 12 + 6 + y] * b[y];
            out[j0 * LB + c] = s;
        }
        rsync<NT>();
        if (k > 0) ric_step_P<LB, NT, 0>(p, r, k, 0);
    }
}

}  // namespace bmpc
