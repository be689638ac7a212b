// Lane-per-robot MPC tick (second generation): ONE THREAD solves one robot, 32 robots per warp share every instruction.
//
// The linear algebra is the stage-wise (Riccati) form of the contact-reduced QP: the problem is an LQR problem
//   X_v = A_v X_{v-1} + B_v u_v + c_v   (MPC.py:165-184, 206-214)
// over VIRTUAL stages v = one stance foot-stage block of 5 free inputs [fx fy fz my mz] (mx is pinned by tau_max[0] =
// tau_min[0], MPC.py:47); a stage with two stance feet is two virtual stages (Z = A X + B_0 u_0, then X' = Z + B_1 u_1).
// A backward sweep over the 12x12 cost-to-go factors  M = Hc + blockdiag(C' D C)  in O(h); Hc is never formed.
//
// What bounds a kernel of this shape is the per-robot state (32 robots per warp), so everything here is built around
// keeping that state small and touching it as few times as possible:
//   * the inequality rows are STRUCTURAL code (box rows touch one component, friction rows two; only the two line-foot
//     rows, MPC.py:253-271, are per-robot vectors, kept in registers): no per-thread matrices, no local memory;
//   * per block only the iterate (u, s, lam), the two step vectors, the stationarity residual and the stage factor
//     (Y = inv(L) F, L) are stored; everything else about the rows (residuals, affine and corrector terms, the step in s
//     and lam) is recomputed from those where it is used, with one fast reciprocal per row;
//   * one interior-point iteration is FOUR sweeps over the block records, in alternating directions so that the turn-around
//     block is still in cache:  (A) backward: apply the previous step, barrier weights, right-hand side, stage factor and
//     the backward half of the predictor solve, fused;  (B) forward half of the predictor + affine step statistics;
//     (C) backward: corrector right-hand side + backward half of its solve;  (D) forward half + step length, the
//     complementarity after the step and the convergence test;
//   * the 12x12 cost-to-go (packed lower triangle, 78 doubles) lives in SHARED memory, lane-interleaved (conflict-free
//     LDS.64); the record of block v of the 32 robots of a warp is one contiguous 41 KB panel of the global workspace
//     (element i of lane l at ws[(v*REC + i)*32 + l]), so every access is a fully used 256-byte line pair.
// Same algorithm otherwise as the warp-per-robot kernel (Mehrotra predictor-corrector to a loose target, active-set
// polish with the SAME per-block null-space / multiplier functions of bmpc_polish.cuh, KKT certificate), written as plain
// scalar C++ so that the identical source is unit-tested on the CPU (tests/lane_host.cu) against the oracle.
// A robot this path does not certify keeps status 1 and is re-solved by the warp-per-robot kernels (bmpc.cu).
#pragma once
#include "bmpc_kernels.cuh"
#include "bmpc_lane_api.h"
#include "bmpc_polish.cuh"

namespace bmpc {

#ifdef __CUDA_ARCH__
#define BMPC_LS 32
#else
#define BMPC_LS 1
#endif

// address-space hints: the views are plain pointers inside a struct, which hides from the compiler that the workspace is
// global and the cost-to-go shared memory (generic LD/ST instead of LDG/STG and LDS/STS)
#ifdef __CUDA_ARCH__
#define BMPC_ASSUME_SPACES()                  \
    __builtin_assume(__isGlobal(ws.p));       \
    __builtin_assume(__isShared(Ps.p))
#else
#define BMPC_ASSUME_SPACES()
#endif

// compiler-only barrier: stops the compiler from keeping shared-memory values (the cost-to-go) in registers from one phase
// of a stage to the next, which costs more in spills than the re-read
#ifdef __CUDA_ARCH__
#define BMPC_CBAR() asm volatile("" ::: "memory")
#else
#define BMPC_CBAR()
#endif

// LaneSolver::run is forced inline in the kernels (see there); the host test build leaves the decision to the compiler (forced,
// gcc needs a quarter of an hour for the twelve instantiations)
#ifdef BMPC_LANE_HOST_ONLY
#define BMPC_RUN_INLINE
#else
#define BMPC_RUN_INLINE __forceinline__
#endif

struct SV {  // strided view of one lane's slice of a lane-interleaved array
    double* p;
    BMPC_HD __forceinline__ double& operator[](int i) const { return p[(size_t)i * BMPC_LS]; }
    BMPC_HD __forceinline__ SV operator+(int o) const { return SV{p + (size_t)o * BMPC_LS}; }
};

// 1/x for the barrier weights: hardware seed + two Newton steps (no special-case branch; x is a positive normal number)
BMPC_HD __forceinline__ double rcp_nr(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#else
    return 1.0 / x;
#endif
}

// candidate inequality rows of one block in storage order (the presolve's order, bmpc_presolve.h): lower bounds of the five
// free components, upper bounds, the four friction rows (MPC.py:220-229), the two line-foot rows (MPC.py:253-271).
// Bit i of DevParams-derived `rowmask` says whether candidate i survived the presolve.
template <int C> struct RLo {};
template <int C> struct RHi {};
template <int R> struct RFr {};
template <int A> struct RLn {};

template <int HZ, int NF>
struct LaneRec {  // per-block record, in doubles
    static constexpr int LB = 5, S = HZ * NF, NR = 16;
    static constexpr int o_B3 = 0;            // 3 x 5: omega rows of the input map (dt Iw^-1 [skew(r) | I] columns of the free components)
    static constexpr int o_ri = 15;           // 6: cz/cy, sz/cy, sz, cz, cz sy/cy, sz sy/cy of the stage (first block of a stage)
    static constexpr int o_c = 21;            // 3: affine term of the omega rows (pinned component)
    static constexpr int o_u = 24;            // 5: iterate / polished solution
    static constexpr int o_xv = 29;           // 5: predictor step (polish: reduced right-hand side / solution)
    static constexpr int o_l = 34;            // 16: multipliers
    static constexpr int o_du = 50;           // 5: corrector / total step
    static constexpr int o_rd = 55;           // 5: stationarity residual
    static constexpr int o_s = 60;            // 16: slacks
    static constexpr int o_Nn = o_du;         // 25: polish: null-space basis (aliases du, rd, s: dead by then)
    static constexpr int o_Y = 76;            // 5 x 12: inv(L) F   (interior point: stored as float, see FT below)
    static constexpr int o_E = o_Y;           // 12: Q (X - xref) between the two halves of a gradient evaluation (aliases Y: no factor is live then)
    static constexpr int o_Lc = 136;          // 15: Cholesky factor of the stage's G, reciprocal diagonal
    static constexpr int hot = 152;           // what the interior-point sweeps touch: [0, hot)
    static constexpr int o_tv = 152;          // 5: gradient scratch (polish), diag(Hc)
    static constexpr int o_pp = 157;          // 5: polish: particular solution of the active rows
    static constexpr int o_am = 162;          // active-row mask of the block (polish)
    static constexpr int o_dim = 163;         // null-space dimension of the block (polish)
    static constexpr int REC = 164;
    static constexpr int total = S * REC;
    static constexpr int smem_doubles = 78;   // per lane: packed cost-to-go
    static constexpr int defer_floats = S * (NR + 1) + 4;  // parked robot (LaneDefer), polish store: per block the active-row mask and NR multipliers; gs, rdmax, iterations
    static constexpr int ipm_floats = S * (5 + 2 * NR) + 4;  // interior-point store: per block u, NR slacks, NR multipliers; gs, iterations
};

// F32: the interior point stores its stage factors (Y, L) as float (the Riccati recursion itself, the cost-to-go and the
// Cholesky stay in double; the stored factor is only used by the two solves of the iteration it belongs to, whose
// direction error ~1e-7 relative an interior point tolerates).  The polish always stores double.
// RM: the surviving candidate rows as a compile-time mask (0 = read the mask from the parameters at run time).  With the
// row set known to the compiler a block's rows are straight-line code in one basic block: the loads and the reciprocal
// chains of different rows overlap.  The two row sets of the reference's limit structure are instantiated:
constexpr unsigned kRowsRef = 0xCF9Bu;  // f_min = 0 (MPC.py:46): fx, fy, my, mz >= lo; fz, my, mz <= hi; two friction rows; line foot
constexpr unsigned kRowsSym = 0xFF98u;  // f_min = -f_max (closed-loop workload): my, mz >= lo; fz, my, mz <= hi; four friction rows; line foot
template <int HZ, int NF, bool F32 = true, unsigned RM = 0u>
struct LaneSolver {
    using L = LaneRec<HZ, NF>;
    static constexpr int LB = 5, S = L::S, NR = L::NR;
    const DevParams& p;
    SV ws;  // global workspace slice of this lane
    SV Ps;  // packed cost-to-go (shared memory on the device)
    int lane_id;
    const double* xfb;  // this robot's feedback state (re-read where it is needed instead of living in 24 registers)
    double ln[2][LB], lnb[2];  // the two line-foot rows in block coordinates and their right-hand sides
    unsigned footbits, rowmask;
    double dt, vm;
    // later passes (bmpc_lane_api.h); mode: 0 first pass, 1 polish pass, 2 interior-point pass; park: this pass parks its stragglers.
    // (A reference to the kernel parameter, like p: read from the constant bank where it is used instead of living in registers.)
    const LaneDefer& df;
    int mode = 0;
    bool park = false;

    BMPC_HD LaneSolver(const DevParams& pp, SV w, SV ps, int lane, const LaneDefer& d) : p(pp), ws(w), Ps(ps), lane_id(lane), df(d) {}
    // element e of parked-robot record q
    BMPC_HD __forceinline__ float& parked(int q, int e) const {
        return df.buf[((size_t)(q / 32) * L::defer_floats + e) * 32 + (q % 32)];
    }
    BMPC_HD __forceinline__ float& parked_ipm(int q, int e) const {
        return df.ipm_buf[((size_t)(q / 32) * L::ipm_floats + e) * 32 + (q % 32)];
    }

    // Bulk prefetch into L2 of the first `ndoubles` entries of block record v of this warp's 32 robots (one contiguous
    // ndoubles x 256-byte panel): issued one stage ahead, so that the sweeps find their operands in L2 instead of paying the
    // DRAM latency on every first touch (the working set of the resident warps is several times the L2).
    BMPC_HD __forceinline__ void prefetch_rec(int v, int ndoubles) const {
#ifdef __CUDA_ARCH__
        if (p.lane_prefetch && v >= 0 && v < S && lane_id == __ffs(__activemask()) - 1) {
            const double* a = ws.p - lane_id + (size_t)v * L::REC * 32;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(ndoubles * 256) : "memory");
        }
#endif
    }
    static constexpr int kFacDoubles = F32 ? 38 : 75;  // extent of the stored stage factor, in doubles

    // stage factor storage: [Y (5 rows of 12) | L (15)] of record r.  As float the 76 values (one of padding) are 19 groups of four,
    // group g of lane l at float4 index g * 32 + l of the record's factor region: one 16-byte access per lane and group (a warp
    // moves 512 contiguous bytes per instruction) instead of four 4-byte ones - the sweeps are bound by the number of memory
    // instructions they issue, not by the bytes.  As double (polish) element i sits at o_Y + i like everything else.
    struct alignas(16) F4 {
        float x, y, z, w;
    };
    BMPC_HD __forceinline__ F4* fac4(SV r, int g) const {
        return reinterpret_cast<F4*>(r.p + (size_t)L::o_Y * BMPC_LS - lane_id) + ((size_t)g * BMPC_LS + lane_id);
    }
    // row a of Y (12 values)
    template <bool F> BMPC_HD __forceinline__ void st_fac_row(SV r, int a, const double (&y)[12]) const {
        if constexpr (F) {
#pragma unroll
            for (int q = 0; q < 3; ++q) *fac4(r, a * 3 + q) = F4{(float)y[4 * q], (float)y[4 * q + 1], (float)y[4 * q + 2], (float)y[4 * q + 3]};
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) r[L::o_Y + a * 12 + j] = y[j];
        }
    }
    template <bool F> BMPC_HD __forceinline__ void ld_fac_row(SV r, int a, double (&y)[12]) const {
        if constexpr (F) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const F4 t = *fac4(r, a * 3 + q);
                y[4 * q] = (double)t.x, y[4 * q + 1] = (double)t.y, y[4 * q + 2] = (double)t.z, y[4 * q + 3] = (double)t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) y[j] = r[L::o_Y + a * 12 + j];
        }
    }
    // the packed Cholesky factor (15 values, reciprocal diagonal)
    template <bool F> BMPC_HD __forceinline__ void st_fac_L(SV r, const double (&Lc)[15]) const {
        if constexpr (F) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *fac4(r, 15 + q) = F4{(float)Lc[4 * q], (float)Lc[4 * q + 1], (float)Lc[4 * q + 2], q < 3 ? (float)Lc[4 * q + 3] : 0.f};
        } else {
#pragma unroll
            for (int e = 0; e < 15; ++e) r[L::o_Y + 60 + e] = Lc[e];
        }
    }
    template <bool F> BMPC_HD __forceinline__ void ld_fac_L(SV r, double (&Lc)[15]) const {
        if constexpr (F) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const F4 t = *fac4(r, 15 + q);
                Lc[4 * q] = (double)t.x, Lc[4 * q + 1] = (double)t.y, Lc[4 * q + 2] = (double)t.z;
                if (q < 3) Lc[4 * q + 3] = (double)t.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 15; ++e) Lc[e] = r[L::o_Y + 60 + e];
        }
    }

    static BMPC_HD __forceinline__ constexpr int pk(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }
    static BMPC_HD __forceinline__ constexpr int tri(int a, int b) { return a * (a + 1) / 2 + b; }  // a >= b
    static BMPC_HD __forceinline__ constexpr int comp_of(int c) { return c < 3 ? c : c + 1; }
    BMPC_HD __forceinline__ SV rec(int v) const { return ws + v * L::REC; }
    BMPC_HD __forceinline__ int foot_of(int v) const { return NF == 2 ? (v & 1) : (int)((footbits >> v) & 1u); }
    BMPC_HD __forceinline__ bool dyn_of(int v) const { return NF == 1 || (v & 1) == 0; }
    BMPC_HD __forceinline__ double Rw(int l, int c) const {  // input weight of free component c of foot l (MPC.py:28, 281)
        const int i0 = (c < 3) ? c : (4 + c);
        return l ? p.R[i0 + 3] : p.R[i0];
    }

    // ---- structural rows ------------------------------------------------------------------------------------------
    template <int C> BMPC_HD __forceinline__ double rdot(RLo<C>, const double (&v)[LB]) const { return -v[C]; }
    template <int C> BMPC_HD __forceinline__ double rdot(RHi<C>, const double (&v)[LB]) const { return v[C]; }
    template <int R> BMPC_HD __forceinline__ double rdot(RFr<R>, const double (&v)[LB]) const {
        return (R < 2 ? v[R & 1] : -v[R & 1]) - p.mu * v[2];
    }
    template <int A> BMPC_HD __forceinline__ double rdot(RLn<A>, const double (&v)[LB]) const {
        return ln[A][0] * v[0] + ln[A][1] * v[1] + ln[A][2] * v[2] + ln[A][3] * v[3] + ln[A][4] * v[4];
    }
    template <int C> BMPC_HD __forceinline__ void radd(RLo<C>, double (&a)[LB], double w) const { a[C] -= w; }
    template <int C> BMPC_HD __forceinline__ void radd(RHi<C>, double (&a)[LB], double w) const { a[C] += w; }
    template <int R> BMPC_HD __forceinline__ void radd(RFr<R>, double (&a)[LB], double w) const {
        a[R & 1] += (R < 2 ? w : -w);
        a[2] -= p.mu * w;
    }
    template <int A> BMPC_HD __forceinline__ void radd(RLn<A>, double (&a)[LB], double w) const {
#pragma unroll
        for (int c = 0; c < LB; ++c) a[c] += ln[A][c] * w;
    }
    template <int C> BMPC_HD __forceinline__ void rrank(RLo<C>, double (&G)[15], double d) const { G[tri(C, C)] += d; }
    template <int C> BMPC_HD __forceinline__ void rrank(RHi<C>, double (&G)[15], double d) const { G[tri(C, C)] += d; }
    template <int R> BMPC_HD __forceinline__ void rrank(RFr<R>, double (&G)[15], double d) const {
        const double md = p.mu * d;
        G[tri(R & 1, R & 1)] += d;
        G[tri(2, R & 1)] += (R < 2 ? -md : md);
        G[tri(2, 2)] += p.mu * md;
    }
    template <int A> BMPC_HD __forceinline__ void rrank(RLn<A>, double (&G)[15], double d) const {
#pragma unroll
        for (int a = 0; a < LB; ++a) {
            const double t = d * ln[A][a];
#pragma unroll
            for (int b = 0; b <= a; ++b) G[tri(a, b)] += t * ln[A][b];
        }
    }
    template <int C> BMPC_HD __forceinline__ double rrhs(RLo<C>) const { return -p.lo6[comp_of(C)]; }
    template <int C> BMPC_HD __forceinline__ double rrhs(RHi<C>) const { return p.hi6[comp_of(C)]; }
    template <int R> BMPC_HD __forceinline__ double rrhs(RFr<R>) const { return 0.0; }
    template <int A> BMPC_HD __forceinline__ double rrhs(RLn<A>) const { return lnb[A]; }
    template <class T> BMPC_HD __forceinline__ void rcoef(T t, double (&c5)[LB]) const {
#pragma unroll
        for (int c = 0; c < LB; ++c) c5[c] = 0.0;
        radd(t, c5, 1.0);
    }

    // f(tag, slot, k): slot = candidate index (compile-time storage position), k = index among the surviving rows
    template <class F> BMPC_HD __forceinline__ void for_rows(F&& f) const {
        int k = 0;
#define BMPC_ROW(bit, tag)      \
    if (RM ? ((RM >> (bit)) & 1u) != 0u : ((rowmask >> (bit)) & 1u) != 0u) {    \
        f(tag{}, (bit), k);     \
        ++k;                    \
    }
        BMPC_ROW(0, RLo<0>) BMPC_ROW(1, RLo<1>) BMPC_ROW(2, RLo<2>) BMPC_ROW(3, RLo<3>) BMPC_ROW(4, RLo<4>)
        BMPC_ROW(5, RHi<0>) BMPC_ROW(6, RHi<1>) BMPC_ROW(7, RHi<2>) BMPC_ROW(8, RHi<3>) BMPC_ROW(9, RHi<4>)
        BMPC_ROW(10, RFr<0>) BMPC_ROW(11, RFr<1>) BMPC_ROW(12, RFr<2>) BMPC_ROW(13, RFr<3>)
        BMPC_ROW(14, RLn<0>) BMPC_ROW(15, RLn<1>)
#undef BMPC_ROW
    }

    // everything the sweeps recompute about one row from the stored iterate and steps
    struct RowStep {
        double is, d, rp, dsa, dla, wc;
    };
    static BMPC_HD __forceinline__ RowStep row_affine(double cu, double cx, double s, double lm, double b) {
        RowStep r;
        r.is = rcp_nr(s);
        r.d = lm * r.is;
        r.rp = cu + s - b;
        r.dsa = -r.rp - cx;
        r.dla = -lm - r.d * r.dsa;
        r.wc = 0.0;
        return r;
    }
    // -dv / v in FP32: only a step LENGTH, cut by step_frac afterwards; NaN/Inf propagate
    static BMPC_HD __forceinline__ float sratio(double dv, double v) {
#ifdef __CUDA_ARCH__
        return __fdividef(-(float)dv, (float)v);  // (no slow-path branch: the rows of a block stay in one basic block)
#else
        return -(float)dv / (float)v;
#endif
    }
    // the stored slacks and multipliers of a block, loaded up front so that their latency overlaps instead of adding up row by row
    BMPC_HD __forceinline__ void load_rows(SV r, double (&sv)[NR], double (&lv)[NR]) const {
        for_rows([&](auto, int slot, int) {
            sv[slot] = r[L::o_s + slot];
            lv[slot] = r[L::o_l + slot];
        });
    }

    // ---- structural input map: column c of B_v is [B3[0..2][c]; vm e_c (c < 3)] on the (omega, v) rows ----------------
    BMPC_HD __forceinline__ void load_B3(SV r, double (&B3)[3][LB]) const {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < LB; ++c) B3[k][c] = r[L::o_B3 + k * LB + c];
    }
    // out[c] = sum_k B[k][c] x6[k]   (B' x, x on the omega/v rows)
    BMPC_HD __forceinline__ void Bt_mul(const double (&B3)[3][LB], const double* x6, double (&out)[LB]) const {
#pragma unroll
        for (int c = 0; c < LB; ++c) {
            double a = B3[0][c] * x6[0] + B3[1][c] * x6[1] + B3[2][c] * x6[2];
            if (c < 3) a += vm * x6[3 + c];
            out[c] = a;
        }
    }
    // x6 += B u
    BMPC_HD __forceinline__ void B_mul_add(const double (&B3)[3][LB], const double (&u)[LB], double* x6) const {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double a = x6[k];
#pragma unroll
            for (int c = 0; c < LB; ++c) a += B3[k][c] * u[c];
            x6[k] = a;
            x6[3 + k] += vm * u[k];
        }
    }
    // rinv9 = dte * [[r0 r1 0] [-sz cz 0] [r6 r7 1]]  (closed form of the inverse at MPC.py:160-164)
    BMPC_HD __forceinline__ void load_r9(SV r, double dte, double (&r9)[9]) const {
        if (NF == 1 || dte != 0.0) {  // (the second block of a stage has A = I and no stored Rinv)
            r9[0] = dte * r[L::o_ri + 0], r9[1] = dte * r[L::o_ri + 1], r9[2] = 0.0;
            r9[3] = -dte * r[L::o_ri + 2], r9[4] = dte * r[L::o_ri + 3], r9[5] = 0.0;
            r9[6] = dte * r[L::o_ri + 4], r9[7] = dte * r[L::o_ri + 5], r9[8] = dte;
        } else {
#pragma unroll
            for (int a = 0; a < 9; ++a) r9[a] = 0.0;
        }
    }
    // z <- A z  (positions += D velocities),  pv <- A' pv  (velocities += D' positions); r9 already carries dt
    static BMPC_HD __forceinline__ void A_mul(const double (&r9)[9], double dte, double (&z)[12]) {
        z[0] += r9[0] * z[6] + r9[1] * z[7];
        z[1] += r9[3] * z[6] + r9[4] * z[7];
        z[2] += r9[6] * z[6] + r9[7] * z[7] + r9[8] * z[8];
#pragma unroll
        for (int a = 0; a < 3; ++a) z[3 + a] += dte * z[9 + a];
    }
    static BMPC_HD __forceinline__ void At_mul(const double (&r9)[9], double dte, double (&pv)[12]) {
        pv[6] += r9[0] * pv[0] + r9[3] * pv[1] + r9[6] * pv[2];
        pv[7] += r9[1] * pv[0] + r9[4] * pv[1] + r9[7] * pv[2];
        pv[8] += r9[8] * pv[2];
#pragma unroll
        for (int a = 0; a < 3; ++a) pv[9 + a] += dte * pv[3 + a];
    }
    // state reference of stage k (MPC.py:61-70), recomputed where it is needed
    BMPC_HD __forceinline__ double xref(int k, int i) const {
        if (k == 0) return xfb[i];
        if (i < 6 && p.x_cmd[i + 6] != 0.0) return xfb[i] + p.x_cmd[i + 6] * (k * dt);
        return p.x_cmd[i];
    }

    // ---- objective gradient  Hc u + g  by a rollout and an adjoint sweep; optionally writes the states -------------
    // u at record offset ou, result at offset oo
    BMPC_HD __forceinline__ void grad(int ou, int oo, double* states) {
        BMPC_ASSUME_SPACES();
        double z[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) z[a] = xfb[a];
#pragma unroll 1
        for (int i = 0; i < HZ; ++i) {
            SV r0 = rec(i * NF);
            double acc[6];
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[k] = r0[L::o_c + k], acc[3 + k] = 0.0;
            acc[5] -= p.g * dt;
#pragma unroll
            for (int li = 0; li < NF; ++li) {
                SV r = rec(i * NF + li);
                double B3[3][LB], ui[LB];
                load_B3(r, B3);
#pragma unroll
                for (int c = 0; c < LB; ++c) ui[c] = r[ou + c];
                B_mul_add(B3, ui, acc);
            }
            double r9[9];
            load_r9(r0, dt, r9);
            A_mul(r9, dt, z);
#pragma unroll
            for (int k = 0; k < 6; ++k) z[6 + k] += acc[k];
#pragma unroll
            for (int a = 0; a < 12; ++a) r0[L::o_E + a] = p.Q[a] * (z[a] - xref(i, a));
            if (states) {
#pragma unroll
                for (int a = 0; a < 12; ++a) states[13 * i + a] = z[a];
                states[13 * i + 12] = 1.0;
            }
        }
        double lam[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) lam[a] = 0.0;
#pragma unroll 1
        for (int i = HZ - 1; i >= 0; --i) {
            SV r0 = rec(i * NF);
#pragma unroll
            for (int a = 0; a < 12; ++a) lam[a] += r0[L::o_E + a];
#pragma unroll
            for (int li = 0; li < NF; ++li) {
                SV r = rec(i * NF + li);
                const int l = foot_of(i * NF + li);
                double B3[3][LB], g5[LB];
                load_B3(r, B3);
                Bt_mul(B3, lam + 6, g5);
#pragma unroll
                for (int c = 0; c < LB; ++c) r[oo + c] = Rw(l, c) * r[ou + c] + g5[c];
            }
            double r9[9];
            load_r9(r0, dt, r9);
            At_mul(r9, dt, lam);
        }
    }

    // P <- A' P A on the packed triangle.  A = [I D; 0 I], D = dt [Rinv 0; 0 I]:  P21 += D' P11,  P22 += D' P12new + P21old D
    BMPC_HD __forceinline__ void congruence_pk(const double (&r9)[9], double dte) {
        SV P = Ps;
        double p11[21], n21[6][6];
#pragma unroll
        for (int e = 0; e < 21; ++e) p11[e] = P[e];
        // row by row, so that a row of the old P21 dies as soon as its row of P22 is done
#pragma unroll
        for (int kp = 0; kp < 6; ++kp) {
            double o[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                o[c] = P[pk(6 + kp, c)];
                if (kp < 3) n21[kp][c] = o[c] + r9[kp] * p11[pk(0, c)] + r9[3 + kp] * p11[pk(1, c)] + r9[6 + kp] * p11[pk(2, c)];
                else n21[kp][c] = o[c] + dte * p11[pk(kp, c)];
                P[pk(6 + kp, c)] = n21[kp][c];
            }
#pragma unroll
            for (int k = 0; k <= kp; ++k) {
                double t1, t2;
                if (kp < 3) t1 = r9[kp] * n21[k][0] + r9[3 + kp] * n21[k][1] + r9[6 + kp] * n21[k][2];
                else t1 = dte * n21[k][kp];
                if (k < 3) t2 = o[0] * r9[k] + o[1] * r9[3 + k] + o[2] * r9[6 + k];
                else t2 = dte * o[k];
                P[pk(6 + kp, 6 + k)] += t1 + t2;
            }
        }
    }

    // t <- N' t (first `dim` components; the others 0)
    static BMPC_HD __forceinline__ void Nt_mul(const double (&N)[LB * LB], int dim, double (&t)[LB]) {
        double o[LB];
#pragma unroll
        for (int a = 0; a < LB; ++a) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < LB; ++c) acc += N[c * LB + a] * t[c];
            o[a] = (a < dim) ? acc : 0.0;
        }
#pragma unroll
        for (int a = 0; a < LB; ++a) t[a] = o[a];
    }

    // ---- one backward stage: stage factor (Y, L stored; cost-to-go updated) + the backward half of a solve ------------
    // G15 holds the stage's input weights on entry; rhs the block's right-hand side.  POL: inputs are w with u = N w.
    // pv is the backward-solve vector carried from stage to stage; w (the intermediate of the solve) is stored at ow.
    template <bool POL>
    BMPC_HD __forceinline__ bool factor_stage(int v, SV r, double (&G)[15], const double (&rhs)[LB], double (&pv)[12], int ow,
                                              const double (&N)[LB * LB], int dim) {
        constexpr bool FS = F32 && !POL;
        SV P = Ps;
        const bool dyn = dyn_of(v);
        const double dte = dyn ? dt : 0.0;
        double B3[3][LB], lo[6][LB];
        load_B3(r, B3);
        // rows 6..11 of P B
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            double pr[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) pr[k] = P[pk(6 + q, 6 + k)];
            Bt_mul(B3, pr, lo[q]);
        }
        BMPC_CBAR();
        // G += B' (P B)[6:12]
#pragma unroll
        for (int a = 0; a < LB; ++a)
#pragma unroll
            for (int j = 0; j <= a; ++j) {
                double acc = G[tri(a, j)] + B3[0][a] * lo[0][j] + B3[1][a] * lo[1][j] + B3[2][a] * lo[2][j];
                if (a < 3) acc += vm * lo[3 + a][j];
                G[tri(a, j)] = acc;
            }
        if constexpr (POL) {  // G <- N' G N, identity on the padding
            double T[LB][LB];
#pragma unroll
            for (int a = 0; a < LB; ++a)
#pragma unroll
                for (int b = 0; b < LB; ++b) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < LB; ++c) acc += G[tri(a > c ? a : c, a > c ? c : a)] * N[c * LB + b];
                    T[a][b] = acc;
                }
#pragma unroll
            for (int a = 0; a < LB; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < LB; ++c) acc += N[c * LB + a] * T[c][b];
                    G[tri(a, b)] = (a < dim) ? acc : ((a == b) ? 1.0 : 0.0);
                }
#pragma unroll
            for (int q = 0; q < 6; ++q) Nt_mul(N, dim, lo[q]);
        }
        // Cholesky in registers (reciprocal diagonal)
        double Lr[LB][LB];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < LB; ++j)
#pragma unroll
            for (int a = j; a < LB; ++a) {
                double x = G[tri(a, j)];
#pragma unroll
                for (int k = 0; k < j; ++k) x -= Lr[a][k] * Lr[j][k];
                if (a == j) {
                    ok = ok && (x > 0.0) && (x < 1e300);
                    Lr[j][j] = 1.0 / sqrt(x);
                } else {
                    Lr[a][j] = x * Lr[j][j];
                }
            }
        if (!ok) return false;
        {
            double Lc[15];
#pragma unroll
            for (int a = 0; a < LB; ++a)
#pragma unroll
                for (int k = 0; k <= a; ++k) Lc[tri(a, k)] = Lr[a][k];
            st_fac_L<FS>(r, Lc);
        }
        // backward half of the solve: w = inv(L) (B' pv - rhs)
        double w[LB];
        {
            double g5[LB];
            Bt_mul(B3, pv + 6, g5);
            if constexpr (POL) Nt_mul(N, dim, g5);
#pragma unroll
            for (int c = 0; c < LB; ++c) {
                double g = g5[c] - rhs[c];
#pragma unroll
                for (int k = 0; k < c; ++k) g -= Lr[c][k] * w[k];
                w[c] = g * Lr[c][c];
                r[ow + c] = w[c];
            }
        }
        // Yt = inv(L) (P B)' (before A): low half in place, P22 -= Yl' Yl
#pragma unroll
        for (int q = 0; q < 6; ++q)
#pragma unroll
            for (int a = 0; a < LB; ++a) {
                double x = lo[q][a];
#pragma unroll
                for (int k = 0; k < a; ++k) x -= Lr[a][k] * lo[q][k];
                lo[q][a] = x * Lr[a][a];
            }
        const bool last = (v == 0);
        if (!last) {
#pragma unroll
            for (int q = 0; q < 6; ++q)
#pragma unroll
                for (int k = 0; k <= q; ++k) {
                    double acc = P[pk(6 + q, 6 + k)];
#pragma unroll
                    for (int a = 0; a < LB; ++a) acc -= lo[q][a] * lo[k][a];
                    P[pk(6 + q, 6 + k)] = acc;
                }
        }
        BMPC_CBAR();
        // high half, one state column at a time: row j of P B -> column j of Yt; P21, P11 downdates; Y = Yt A
        double hi[6][LB];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double pr[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) pr[k] = P[pk(6 + k, j)];
            Bt_mul(B3, pr, hi[j]);
            if constexpr (POL) Nt_mul(N, dim, hi[j]);
#pragma unroll
            for (int a = 0; a < LB; ++a) {
                double x = hi[j][a];
#pragma unroll
                for (int k = 0; k < a; ++k) x -= Lr[a][k] * hi[j][k];
                hi[j][a] = x * Lr[a][a];
            }
            if (!last) {
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    double acc = P[pk(6 + k, j)];
#pragma unroll
                    for (int a = 0; a < LB; ++a) acc -= lo[k][a] * hi[j][a];
                    P[pk(6 + k, j)] = acc;
                }
#pragma unroll
                for (int k = 0; k <= j; ++k) {
                    double acc = P[pk(j, k)];
#pragma unroll
                    for (int a = 0; a < LB; ++a) acc -= hi[j][a] * hi[k][a];
                    P[pk(j, k)] = acc;
                }
            }
            BMPC_CBAR();
        }
        // Y = Yt A: columns 6..11 += columns 0..5 times D; store; pv <- A' pv - Y' w
        double r9[9];
        load_r9(r, dte, r9);
#pragma unroll
        for (int a = 0; a < LB; ++a) {
            lo[0][a] += hi[0][a] * r9[0] + hi[1][a] * r9[3] + hi[2][a] * r9[6];
            lo[1][a] += hi[0][a] * r9[1] + hi[1][a] * r9[4] + hi[2][a] * r9[7];
            lo[2][a] += hi[2][a] * r9[8];
#pragma unroll
            for (int k = 3; k < 6; ++k) lo[k][a] += dte * hi[k][a];
        }
        At_mul(r9, dte, pv);
#pragma unroll
        for (int a = 0; a < LB; ++a) {
            double y[12];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                y[j] = hi[j][a], y[6 + j] = lo[j][a];
                pv[j] -= hi[j][a] * w[a];
                pv[6 + j] -= lo[j][a] * w[a];
            }
            st_fac_row<FS>(r, a, y);
        }
        if (last) return true;
        BMPC_CBAR();
        // cost-to-go of the previous state: A' P A (+ Q)
        if (dyn) {
            congruence_pk(r9, dte);
#pragma unroll
            for (int a = 0; a < 12; ++a) P[pk(a, a)] += p.Q[a];
        }
        return true;
    }

    BMPC_HD __forceinline__ void init_P() {
        SV P = Ps;
#pragma unroll 1
        for (int e = 0; e < 78; ++e) P[e] = 0.0;
#pragma unroll
        for (int a = 0; a < 12; ++a) P[pk(a, a)] = p.Q[a];
    }

    // backward half of a solve with the stored factor (one stage): w = inv(L)(B' pv - rhs) stored at ow; pv <- A' pv - Y' w
    template <bool POL>
    BMPC_HD __forceinline__ void solve_back_stage(int v, SV r, const double (&rhs)[LB], double (&pv)[12], int ow,
                                                  const double (&N)[LB * LB], int dim) {
        constexpr bool FS = F32 && !POL;
        const double dte = dyn_of(v) ? dt : 0.0;
        double B3[3][LB], g5[LB], w[LB], Lc[15], r9[9];
        load_B3(r, B3);
        load_r9(r, dte, r9);
        ld_fac_L<FS>(r, Lc);
        Bt_mul(B3, pv + 6, g5);
        if constexpr (POL) Nt_mul(N, dim, g5);
#pragma unroll
        for (int c = 0; c < LB; ++c) {
            double g = g5[c] - rhs[c];
#pragma unroll
            for (int k = 0; k < c; ++k) g -= Lc[tri(c, k)] * w[k];
            w[c] = g * Lc[tri(c, c)];
        }
        At_mul(r9, dte, pv);
#pragma unroll
        for (int a = 0; a < LB; ++a) {
            double y[12];
            ld_fac_row<FS>(r, a, y);
#pragma unroll
            for (int j = 0; j < 12; ++j) pv[j] -= y[j] * w[a];
        }
        // (stored last: a store in the middle would pin the factor loads behind it)
#pragma unroll
        for (int c = 0; c < LB; ++c) r[ow + c] = w[c];
    }

    // forward half of a solve (one stage): x = -inv(L') (w + Y z); z <- A z + B x (POL: B N x).  Returns x.
    template <bool POL>
    BMPC_HD __forceinline__ void solve_fwd_stage(int v, SV r, int ow, double (&z)[12], double (&xs)[LB]) const {
        constexpr bool FS = F32 && !POL;
        double t[LB], Lc[15];
        ld_fac_L<FS>(r, Lc);
#pragma unroll
        for (int a = 0; a < LB; ++a) {
            double x = r[ow + a], y[12];
            ld_fac_row<FS>(r, a, y);
#pragma unroll
            for (int j = 0; j < 12; ++j) x += y[j] * z[j];
            t[a] = x;
        }
#pragma unroll
        for (int a = LB - 1; a >= 0; --a) {
            double x = t[a];
#pragma unroll
            for (int k = a + 1; k < LB; ++k) x -= Lc[tri(k, a)] * xs[k];
            xs[a] = x * Lc[tri(a, a)];
        }
#pragma unroll
        for (int a = 0; a < LB; ++a) xs[a] = -xs[a];
    }
    BMPC_HD __forceinline__ void advance_z(int v, SV r, const double (&ustep)[LB], double (&z)[12]) const {
        const double dte = dyn_of(v) ? dt : 0.0;
        double B3[3][LB], r9[9];
        load_B3(r, B3);
        load_r9(r, dte, r9);
        A_mul(r9, dte, z);
        B_mul_add(B3, ustep, z + 6);
    }
    // the same with the stage's maps loaded by the caller (at the top of the stage, so that their latency overlaps the solve: the
    // compiler cannot move these loads above the stores of the stage itself)
    BMPC_HD __forceinline__ void advance_z(const double (&B3)[3][LB], const double (&r9)[9], double dte, const double (&ustep)[LB],
                                           double (&z)[12]) const {
        A_mul(r9, dte, z);
        B_mul_add(B3, ustep, z + 6);
    }

    // ---- the row pass of sweep A for one block: (APPLY) the previous iteration's step in u, s, lam and rd, recomputed from the
    //      stored step vectors; then barrier weights, the predictor right-hand side (accumulated into gacc, which enters holding
    //      rd) and the block's input weights G.  Returns the block's complementarity sum. -----------------------------------
    template <bool APPLY>
    BMPC_HD __forceinline__ double rows_A(SV r, double alpha, double tgt, const double (&u)[LB], double (&gacc)[LB], double (&G)[15]) const {
        double sv[NR], lv[NR], x5[LB], d5[LB];
        load_rows(r, sv, lv);
        if constexpr (APPLY) {
#pragma unroll
            for (int c = 0; c < LB; ++c) {
                x5[c] = r[L::o_xv + c];
                d5[c] = r[L::o_du + c];
                gacc[c] *= (1.0 - alpha);
                r[L::o_rd + c] = gacc[c];
                r[L::o_u + c] = u[c] + alpha * d5[c];
            }
        }
        double part = 0.0;
        for_rows([&](auto tag, int slot, int) {
            double s = sv[slot], lm = lv[slot];
            const double b = rrhs(tag);
            double cu = rdot(tag, u);
            if constexpr (APPLY) {
                const double cx = rdot(tag, x5), cd = rdot(tag, d5);
                RowStep q = row_affine(cu, cx, s, lm, b);
                const double wc = (q.dsa * q.dla - tgt) * q.is;
                const double ds = -q.rp - cd, dl = -lm - wc - q.d * ds;
                s += alpha * ds;
                lm += alpha * dl;
                cu += alpha * cd;
                r[L::o_s + slot] = s;
                r[L::o_l + slot] = lm;
            }
            const double d = lm * rcp_nr(s), rp = cu + s - b;
            part += s * lm;
            radd(tag, gacc, d * rp - lm);
            rrank(tag, G, d);
        });
        return part;
    }

    // ---- per-block active-set algebra of the polish, on the structural rows (same mathematics as block_nullspace /
    //      block_dual_fast of bmpc_polish.cuh, which serve the warp-per-robot kernels from a dense row matrix; here the rows are
    //      code, every row is visited by every lane with a 0 / 1 weight, and nothing is indexed at run time, so the 32 robots of
    //      a warp stay on one path and nothing lives in local memory).  Bit k of `mk` = k-th surviving row. ------------------
    BMPC_HD __forceinline__ void active_normal(unsigned mk, const double (&rhs)[LB], double (&A)[LB][LB + 1], int (&perm)[LB], int& np) const {
        double G[15];
#pragma unroll
        for (int e = 0; e < 15; ++e) G[e] = 0.0;
        for_rows([&](auto tag, int, int k) { rrank(tag, G, ((mk >> k) & 1u) ? 1.0 : 0.0); });
#pragma unroll
        for (int a = 0; a < LB; ++a) {
#pragma unroll
            for (int b = 0; b <= a; ++b) A[a][b] = G[tri(a, b)], A[b][a] = G[tri(a, b)];
            A[a][LB] = rhs[a];
        }
        gauss_jordan_diag<LB>(A, perm, np);
    }
    // position space -> component space: out[c] = A[i][LB] of the pivot position i with perm[i] == c, 0 for a free component
    static BMPC_HD __forceinline__ void unpermute(const double (&A)[LB][LB + 1], const int (&perm)[LB], int np, double (&out)[LB]) {
#pragma unroll
        for (int c = 0; c < LB; ++c) {
            double x = 0.0;
#pragma unroll
            for (int i = 0; i < LB; ++i) x = (i < np && perm[i] == c) ? A[i][LB] : x;
            out[c] = x;
        }
    }
    // affine set {x : C_A x = b_A} of the active rows: particular solution -> o_pp, basis -> o_Nn, dimension -> o_dim of the
    // record; false if the rows are inconsistent
    BMPC_HD __forceinline__ bool blk_nullspace(SV r, unsigned mk) const {
        double rhs[LB];
#pragma unroll
        for (int c = 0; c < LB; ++c) rhs[c] = 0.0;
        double bmax = 1.0;
        for_rows([&](auto tag, int, int k) {
            const bool on = ((mk >> k) & 1u) != 0u;
            const double bk = rrhs(tag);
            bmax = on ? fmax(bmax, fabs(bk)) : bmax;
            radd(tag, rhs, on ? bk : 0.0);
        });
        double A[LB][LB + 1];
        int perm[LB], np;
        active_normal(mk, rhs, A, perm, np);
        double p5[LB];
        unpermute(A, perm, np, p5);
#pragma unroll
        for (int c = 0; c < LB; ++c) r[L::o_pp + c] = p5[c];
        // basis: column a = q - np of N for every free position q (the record is this lane's own memory: run-time offsets are fine)
#pragma unroll
        for (int e = 0; e < LB * LB; ++e) r[L::o_Nn + e] = 0.0;
#pragma unroll
        for (int q = 0; q < LB; ++q)
            if (q >= np) {
#pragma unroll
                for (int i = 0; i < LB; ++i) r[L::o_Nn + perm[i] * LB + (q - np)] = (i < np) ? -A[i][q] : ((i == q) ? 1.0 : 0.0);
            }
        r[L::o_dim] = (double)(LB - np);
        bool ok = true;
        for_rows([&](auto tag, int, int k) {
            const bool on = ((mk >> k) & 1u) != 0u;
            if (on && fabs(rdot(tag, p5) - rrhs(tag)) > 1e-7 * bmax) ok = false;
        });
        return ok;
    }
    // fast multiplier check (block_dual_fast): y = lam + C_A z with G z = r - C_A' lam reproduces r = -(gradient, at o_tv)
    // exactly and stays close to the interior-point multipliers (at o_l); true if y >= 0 and C_A' y = r
    BMPC_HD __forceinline__ bool blk_dual_fast(SV r, unsigned mk, double gscale) const {
        double rr[LB], rho[LB], lv[NR];
#pragma unroll
        for (int c = 0; c < LB; ++c) rr[c] = -r[L::o_tv + c], rho[c] = rr[c];
        for_rows([&](auto tag, int slot, int k) {
            lv[slot] = r[L::o_l + slot];
            radd(tag, rho, ((mk >> k) & 1u) ? -lv[slot] : 0.0);
        });
        double A[LB][LB + 1];
        int perm[LB], np;
        active_normal(mk, rho, A, perm, np);
        double z5[LB];
        unpermute(A, perm, np, z5);
        bool ok = true;
        for_rows([&](auto tag, int slot, int k) {
            const bool on = ((mk >> k) & 1u) != 0u;
            const double y = lv[slot] + rdot(tag, z5);
            if (on && !(y >= 0.0)) ok = false;
            radd(tag, rr, on ? -y : 0.0);
        });
        double rmax = 0.0;
#pragma unroll
        for (int c = 0; c < LB; ++c) rmax = fmax(rmax, fabs(rr[c]));
        return ok && rmax <= 1e-9 * gscale;
    }
    // exact multiplier check of one block (rare: only when the fast check is undecided): Lawson-Hanson NNLS of bmpc_polish.cuh
    // on a dense copy of the rows, built here so that the dense copy is only touched on this path
    BMPC_HD __forceinline__ bool blk_dual_exact(SV r, unsigned mk, double gscale, unsigned* drop) const {
        double Cb[NR * LB], rneg[LB];
        for_rows([&](auto tag, int, int k) {
            double c5[LB];
            rcoef(tag, c5);
#pragma unroll
            for (int c = 0; c < LB; ++c) Cb[k * LB + c] = c5[c];
        });
#pragma unroll
        for (int c = 0; c < LB; ++c) rneg[c] = -r[L::o_tv + c];
        return block_dual_check<LB>(Cb, p.mb, mk, rneg, gscale, drop);
    }

    // ---- group synchronisation: the warps of a CTA can run in lockstep (p.lane_sync: 0 none, 1 per iteration / phase, 2 also per
    //      stage of every sweep) so that they fetch the same instructions at the same time.  Every thread of the group takes
    //      the same control path; `act`-style flags, not returns, switch a lane off. ------------------------------------
    BMPC_HD __forceinline__ bool group_any(bool x) const {
#ifdef __CUDA_ARCH__
        if (p.lane_sync & 3) return __syncthreads_or(x) != 0;
        return __any_sync(0xffffffffu, x) != 0;
#else
        return x;
#endif
    }
    // (p.lane_sync & 8: a barrier only before the first stage of a sweep - the warps of a CTA then run the same sweep, i.e. the
    //  same loop body, without waiting for each other at every stage)
    BMPC_HD __forceinline__ void stage_sync(bool first) const {
#ifdef __CUDA_ARCH__
        if ((p.lane_sync & 3) >= 2 && (first || !(p.lane_sync & 8))) __syncthreads();
#endif
    }
    // the polish: in lockstep like the interior point, or (p.lane_sync & 4) every warp on its own - the number of polish rounds
    // differs from robot to robot, and a CTA in lockstep runs the rounds of its slowest robot
    BMPC_HD __forceinline__ bool group_any_polish(bool x) const {
#ifdef __CUDA_ARCH__
        if ((p.lane_sync & 3) && !(p.lane_sync & 4)) return __syncthreads_or(x) != 0;
        return __any_sync(0xffffffffu, x) != 0;
#else
        return x;
#endif
    }
    BMPC_HD __forceinline__ void stage_sync_polish(bool first) const {
#ifdef __CUDA_ARCH__
        if ((p.lane_sync & 3) >= 2 && !(p.lane_sync & 4) && (first || !(p.lane_sync & 8))) __syncthreads();
#endif
    }

    // ---- the whole tick; inst < 0: no robot for this lane (it only takes part in the group's barriers); q: position of the
    //      robot among the parked ones (second pass only) ----------------------------------------------------------------
    //      (forced inline: as a real call the solver object - line-foot rows, pointers, scalars read in every row pass - lives
    //      in local memory; the compiler's own heuristic stops inlining a function of this size: measured +30 % kernel time)
    BMPC_HD BMPC_RUN_INLINE void run(const IoPtrs& io, int inst, int q = -1) {
        BMPC_ASSUME_SPACES();
        dt = p.dt;
        vm = dt / p.mass;
        // which candidate rows survived the presolve (same order as the presolve's list)
        rowmask = 0u;
        for (int k = 0; k < p.mb; ++k) {
            const int kind = p.row_kind[k], arg = p.row_arg[k];
            rowmask |= 1u << (kind == ROW_LO ? arg : kind == ROW_HI ? 5 + arg : kind == ROW_FRIC ? 10 + arg : 14 + arg);
        }
        const int mb = p.mb, m = S * mb;
        const double pin = p.lo6[3];  // value of the pinned component mx
        bool act = inst >= 0;
        int cont0[2] = {0, 0};
        double gs = 1.0, rdmax = 0.0;

        if (act) {
            // ---- 0. inputs; this path takes robots with exactly NF stance feet in every stage and mx as the pinned component ----
            bool bad = p.LB != 5 || p.npinned != 1 || p.pinned[0] != 3 || mb > NR || (RM != 0u && rowmask != RM);
            xfb = io.x_fb + (size_t)inst * 12;
#pragma unroll
            for (int a = 0; a < 12; ++a) bad = bad || !isfinite(xfb[a]);
            double foot[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                foot[a] = io.foot[(size_t)inst * 6 + a];
                bad = bad || !isfinite(foot[a]);
            }
            footbits = 0u;
#pragma unroll 1
            for (int s = 0; s < HZ; ++s) {
                const int c0 = io.contact[(size_t)inst * 2 * HZ + 2 * s] ? 1 : 0, c1 = io.contact[(size_t)inst * 2 * HZ + 2 * s + 1] ? 1 : 0;
                if (s == 0) cont0[0] = c0, cont0[1] = c1;
                if (c0 + c1 != NF) bad = true;
                if (NF == 1 && c1) footbits |= 1u << s;
            }
            // ---- 1. references, per-stage dynamics and input maps (MPC.py:61-109, 148-185) ----
            bool singular = false;
            double ub[LB];
            if (!bad) {
                const int phase_k = io.phase_k[inst];
                double x_fb[12], rotn[9];
#pragma unroll
                for (int a = 0; a < 12; ++a) x_fb[a] = xfb[a];
                const double hh = (double)p.h;
                const double ex = p.kv * (x_fb[3] - p.x_cmd[3]), ey = p.kv * (x_fb[4] - p.x_cmd[4]);
                const double x1 = x_fb[3] + x_fb[9] * 1 / 2 * hh / 2 * dt + ex;
                const double x2 = x_fb[3] + x_fb[9] * 1 / 2 * hh * dt + ex;
                const double y1 = x_fb[4] + x_fb[10] * 1 / 2 * hh / 2 * dt + ey;
                const double y2 = x_fb[10] + x_fb[10] * 1 / 2 * hh * dt + ey;  // MPC.py:87 starts from x_fb[10]
                eul2rotm(x_fb, rotn);
                double u6[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const double lo = p.lo6[c], hi = p.hi6[c];
                    const double v0 = fmin(fmax(0.0, lo + 0.1 * (hi - lo)), hi - 0.1 * (hi - lo));
                    u6[c] = (hi > lo) ? v0 : lo;
                }
                u6[2] = p.lo6[2] + p.init_fz_frac * (p.hi6[2] - p.lo6[2]);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double lo = fmax(p.lo6[c], -p.mu * u6[2]), hi = fmin(p.hi6[c], p.mu * u6[2]);
                    if (p.hi6[c] > p.lo6[c]) u6[c] = 0.5 * (lo + hi);
                }
#pragma unroll
                for (int c = 0; c < LB; ++c) ub[c] = u6[comp_of(c)];
                // the two line-foot rows (MPC.py:253-271) in block coordinates; the pinned component goes to the right-hand side
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const double len = (a == 0) ? p.lh_eff : p.lt_eff;
                    const double sg = (a == 0) ? 1.0 : -1.0;
                    ln[a][0] = -len * rotn[2], ln[a][1] = -len * rotn[5], ln[a][2] = -len * rotn[8];
                    ln[a][3] = sg * rotn[4], ln[a][4] = sg * rotn[7];
                    lnb[a] = 0.0 - sg * rotn[1] * pin;
                }
                const int kk = phase_k % 5;
                const bool one = (cont0[0] + cont0[1] == 1);
#pragma unroll 1
                for (int k = 0; k < HZ; ++k) {
                    int sel = 0;
                    if (one) sel = (k < 5 - kk) ? 0 : ((k < 10 - kk) ? 1 : 2);
                    const double xr0 = xref(k, 0), xr1 = xref(k, 1), xr2 = xref(k, 2), xr3 = xref(k, 3), xr4 = xref(k, 4), xr5 = xref(k, 5);
                    double sz, cz, sy, cy, sx, cx;  // dynamics read x[0] as yaw, x[1] pitch, x[2] roll (MPC.py:151-153)
                    sincos(xr0, &sz, &cz);
                    sincos(xr1, &sy, &cy);
                    sincos(xr2, &sx, &cx);
                    double rot[9];  // Rx(roll) Ry(pitch) Rz(yaw)  (extrinsic 'zyx', MPC.py:156)
                    rot[0] = cy * cz, rot[1] = -cy * sz, rot[2] = sy;
                    rot[3] = sx * sy * cz + cx * sz, rot[4] = -sx * sy * sz + cx * cz, rot[5] = -sx * cy;
                    rot[6] = -cx * sy * cz + sx * sz, rot[7] = cx * sy * sz + sx * cz, rot[8] = cx * cy;
                    double tmp[9], iw[9], ii[9];
                    mat3_mul(p.inertia, rot, tmp);
                    mat3_tmul(rot, tmp, iw);
                    if (!mat3_inv(iw, ii)) singular = true;
                    const double icp = 1.0 / cy;
                    if (!isfinite(icp) || fabs(cy) < 1e-9) singular = true;
                    SV r0 = rec(k * NF);
                    r0[L::o_ri + 0] = cz * icp, r0[L::o_ri + 1] = sz * icp, r0[L::o_ri + 2] = sz, r0[L::o_ri + 3] = cz;
                    r0[L::o_ri + 4] = cz * sy * icp, r0[L::o_ri + 5] = sz * sy * icp;
                    double cw[3] = {0, 0, 0};
#pragma unroll
                    for (int li = 0; li < NF; ++li) {
                        const int v = k * NF + li, l = foot_of(v);
                        SV r = rec(v);
                        // foot reference (MPC.py:72-109): the current feet, then foothold 1, then foothold 2 (both feet at the same x, y; z = 0)
                        const double f0 = sel == 0 ? (l ? foot[3] : foot[0]) : (sel == 1 ? x1 : x2);
                        const double f1 = sel == 0 ? (l ? foot[4] : foot[1]) : (sel == 1 ? y1 : y2);
                        const double f2 = sel == 0 ? (l ? foot[5] : foot[2]) : 0.0;
                        const double q0 = f0 - xr3, q1 = f1 - xr4, q2 = f2 - xr5;
#pragma unroll
                        for (int a = 0; a < 3; ++a) {  // dt Iw^{-1} [skew(r) | I]  (MPC.py:174-179, 184): columns fx fy fz (mx) my mz
                            const double i0 = ii[3 * a], i1 = ii[3 * a + 1], i2 = ii[3 * a + 2];
                            r[L::o_B3 + a * LB + 0] = dt * (i1 * q2 - i2 * q1);
                            r[L::o_B3 + a * LB + 1] = dt * (i2 * q0 - i0 * q2);
                            r[L::o_B3 + a * LB + 2] = dt * (i0 * q1 - i1 * q0);
                            r[L::o_B3 + a * LB + 3] = dt * i1;
                            r[L::o_B3 + a * LB + 4] = dt * i2;
                            cw[a] += dt * i0 * pin;
                        }
                    }
#pragma unroll
                    for (int a = 0; a < 3; ++a) r0[L::o_c + a] = cw[a];
                }
            }
            if (bad || singular) {  // not this path's robot (or bad input): the warp-per-robot kernels take it
                io.status[inst] = 1;
                io.iters[inst] = 0;
                act = false;
            }

            // ---- 2. interior point: start ----
            if (act && mode == 2) {
                // interior-point pass: the parked interior point; its stationarity residual rd = Hc u + g + C' lam is recomputed
                constexpr int PB = 5 + 2 * NR;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
#pragma unroll
                    for (int c = 0; c < LB; ++c) r[L::o_u + c] = (double)parked_ipm(q, v * PB + c);
                    for_rows([&](auto, int slot, int) {
                        r[L::o_s + slot] = (double)parked_ipm(q, v * PB + 5 + slot);
                        r[L::o_l + slot] = (double)parked_ipm(q, v * PB + 5 + NR + slot);
                    });
                }
                gs = (double)parked_ipm(q, S * PB);
                grad(L::o_u, L::o_tv, nullptr);
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
                    double acc[LB];
#pragma unroll
                    for (int c = 0; c < LB; ++c) acc[c] = r[L::o_tv + c];
                    for_rows([&](auto tag, int slot, int) { radd(tag, acc, r[L::o_l + slot]); });
#pragma unroll
                    for (int c = 0; c < LB; ++c) {
                        r[L::o_rd + c] = acc[c];
                        rdmax = fmax(rdmax, fabs(acc[c]));
                    }
                }
            }
            if (act && mode == 0) {
                // g = gradient at u = 0 (scale of the problem)
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
#pragma unroll
                    for (int c = 0; c < LB; ++c) r[L::o_u + c] = 0.0;
                }
                grad(L::o_u, L::o_tv, nullptr);
                gs = 0.0;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
#pragma unroll
                    for (int c = 0; c < LB; ++c) gs = fmax(gs, fabs(r[L::o_tv + c]));
                }
                gs += 1.0;
                // start point: the same interior point in every block
                double part = 0.0;
                double sl0[NR];
                for_rows([&](auto tag, int slot, int) {
                    double sl = rrhs(tag) - rdot(tag, ub);
                    if (!(sl > 1e-3)) sl = 1.0;
                    sl0[slot] = sl;
                    part += sl * (double)S;
                });
                const double mu0 = p.mu0_scale * part / (double)m;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
#pragma unroll
                    for (int c = 0; c < LB; ++c) r[L::o_u + c] = ub[c];
                    for_rows([&](auto, int slot, int) {
                        r[L::o_s + slot] = sl0[slot];
                        r[L::o_l + slot] = mu0 / sl0[slot];
                    });
                }
                grad(L::o_u, L::o_tv, nullptr);
#pragma unroll 1
                for (int v = 0; v < S; ++v) {  // rd = Hc u + g + C' lam
                    SV r = rec(v);
                    double acc[LB];
#pragma unroll
                    for (int c = 0; c < LB; ++c) acc[c] = r[L::o_tv + c];
                    for_rows([&](auto tag, int slot, int) { radd(tag, acc, r[L::o_l + slot]); });
#pragma unroll
                    for (int c = 0; c < LB; ++c) {
                        r[L::o_rd + c] = acc[c];
                        rdmax = fmax(rdmax, fabs(acc[c]));
                    }
                }
            }
        }

        const double mu_target = p.mu_tol * gs;
        int status = 1, it = 0;
        double alpha = 0.0, tgt = 0.0;
        const double dummyN[LB * LB] = {0.0};  // (unused: the interior point works on the inputs themselves)
        bool ipm = act && mode != 1, park_ipm = false;
        if (ipm && mode == 2) it = (int)parked_ipm(q, S * (5 + 2 * NR) + 1);
#pragma unroll 1
        while (group_any(ipm)) {
            if (ipm && it >= p.max_iter) status = 1, ipm = false;
            if (ipm) ++it;
            // -- sweep A (backward): apply the previous step, barrier weights, right-hand side, stage factor, backward half
            //    of the predictor solve
            double pv[12];
            double part = 0.0;
            if (ipm) {
                init_P();
#pragma unroll
                for (int a = 0; a < 12; ++a) pv[a] = 0.0;
            }
#pragma unroll 1
            for (int v = S - 1; v >= 0; --v) {
                stage_sync(v == S - 1);
                if (!ipm) continue;
                prefetch_rec(v - 1, L::o_Y);
                SV r = rec(v);
                const int l = foot_of(v);
                double u[LB], gacc[LB], G[15];
#pragma unroll
                for (int c = 0; c < LB; ++c) {
                    u[c] = r[L::o_u + c];
                    gacc[c] = r[L::o_rd + c];
                }
#pragma unroll
                for (int a = 0; a < LB; ++a)
#pragma unroll
                    for (int b = 0; b <= a; ++b) G[tri(a, b)] = (a == b) ? Rw(l, a) : 0.0;
                if (alpha != 0.0) part += rows_A<true>(r, alpha, tgt, u, gacc, G);
                else part += rows_A<false>(r, alpha, tgt, u, gacc, G);  // first iteration: no step to apply (du is not initialised yet)
                double rhs[LB];
#pragma unroll
                for (int c = 0; c < LB; ++c) rhs[c] = -gacc[c];
                if (!factor_stage<false>(v, r, G, rhs, pv, L::o_xv, dummyN, LB)) status = 2, ipm = false;
            }
            const double mu = part / (double)m;
            // -- sweep B (forward): predictor step and its statistics
            float ratio = 0.f;
            part = 0.0;
            {
                double z[12];
#pragma unroll
                for (int a = 0; a < 12; ++a) z[a] = 0.0;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    stage_sync(v == 0);
                    if (!ipm) continue;
                    prefetch_rec(v + 1, L::o_Y + kFacDoubles);
                    SV r = rec(v);
                    double xs[LB], u[LB], sv[NR], lv[NR];
                    load_rows(r, sv, lv);
#pragma unroll
                    for (int c = 0; c < LB; ++c) u[c] = r[L::o_u + c];
                    const double dte = dyn_of(v) ? dt : 0.0;
                    double B3[3][LB], r9[9];
                    load_B3(r, B3);
                    load_r9(r, dte, r9);
                    solve_fwd_stage<false>(v, r, L::o_xv, z, xs);
#pragma unroll
                    for (int c = 0; c < LB; ++c) r[L::o_xv + c] = xs[c];
                    advance_z(B3, r9, dte, xs, z);
                    for_rows([&](auto tag, int slot, int) {
                        const double s = sv[slot], lm = lv[slot];
                        const RowStep q = row_affine(rdot(tag, u), rdot(tag, xs), s, lm, rrhs(tag));
                        ratio = fmaxf(ratio, fmaxf(sratio(q.dsa, s), sratio(q.dla, lm)));
                        part += q.dsa * q.dla;
                    });
                }
            }
            if (ipm) {
                const double a_aff = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
                const double mu_aff = mu * (1.0 - a_aff) + a_aff * a_aff * part / (double)m;
                double sigma = mu_aff / mu;
                sigma = sigma * sigma * sigma;
                tgt = sigma * mu;
#pragma unroll
                for (int a = 0; a < 12; ++a) pv[a] = 0.0;
            }
            // -- sweep C (backward): corrector right-hand side C' wc, wc = (dsa dla - sigma mu) / s, backward half of its solve
#pragma unroll 1
            for (int v = S - 1; v >= 0; --v) {
                stage_sync(v == S - 1);
                if (!ipm) continue;
                prefetch_rec(v - 1, L::o_Y + kFacDoubles);
                SV r = rec(v);
                double u[LB], x5[LB], gacc[LB], sv[NR], lv[NR];
                load_rows(r, sv, lv);
#pragma unroll
                for (int c = 0; c < LB; ++c) {
                    u[c] = r[L::o_u + c];
                    x5[c] = r[L::o_xv + c];
                    gacc[c] = 0.0;
                }
                for_rows([&](auto tag, int slot, int) {
                    const double s = sv[slot], lm = lv[slot];
                    const RowStep q = row_affine(rdot(tag, u), rdot(tag, x5), s, lm, rrhs(tag));
                    radd(tag, gacc, (q.dsa * q.dla - tgt) * q.is);
                });
                solve_back_stage<false>(v, r, gacc, pv, L::o_du, dummyN, LB);
            }
            // -- sweep D (forward): total step, step length, complementarity after the step
            ratio = 0.f;
            double sum_sl = 0.0, sum_x = 0.0, sum_dd = 0.0;
            {
                double z[12];
#pragma unroll
                for (int a = 0; a < 12; ++a) z[a] = 0.0;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    stage_sync(v == 0);
                    if (!ipm) continue;
                    prefetch_rec(v + 1, L::o_Y + kFacDoubles);
                    SV r = rec(v);
                    double xs[LB], u[LB], x5[LB], d5[LB], sv[NR], lv[NR];
                    load_rows(r, sv, lv);
#pragma unroll
                    for (int c = 0; c < LB; ++c) {
                        u[c] = r[L::o_u + c];
                        x5[c] = r[L::o_xv + c];
                    }
                    const double dte = dyn_of(v) ? dt : 0.0;
                    double B3[3][LB], r9[9];
                    load_B3(r, B3);
                    load_r9(r, dte, r9);
                    solve_fwd_stage<false>(v, r, L::o_du, z, xs);
                    advance_z(B3, r9, dte, xs, z);
#pragma unroll
                    for (int c = 0; c < LB; ++c) {
                        d5[c] = xs[c] + x5[c];
                        r[L::o_du + c] = d5[c];
                    }
                    for_rows([&](auto tag, int slot, int) {
                        const double s = sv[slot], lm = lv[slot];
                        const RowStep q = row_affine(rdot(tag, u), rdot(tag, x5), s, lm, rrhs(tag));
                        const double wc = (q.dsa * q.dla - tgt) * q.is;
                        const double ds = -q.rp - rdot(tag, d5), dl = -lm - wc - q.d * ds;
                        ratio = fmaxf(ratio, fmaxf(sratio(ds, s), sratio(dl, lm)));
                        sum_sl += s * lm;
                        sum_x += s * dl + lm * ds;
                        sum_dd += ds * dl;
                    });
                }
            }
            if (ipm) {
                if (!isfinite(ratio)) {
                    status = 2, ipm = false;
                } else {
                    const double a2 = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
                    alpha = (it > 14 ? 0.9 : p.step_frac) * a2;  // applied by the next sweep A (or by the hand-over below)
                    rdmax *= (1.0 - alpha);
                    const double mu_next = (sum_sl + alpha * (sum_x + alpha * sum_dd)) / (double)m;
                    if (mu_next <= mu_target && rdmax <= p.rd_tol * mu_target) {
                        status = 0;
                        ++it;  // (the iteration count includes the one that only tests convergence, as in the warp-per-robot kernels)
                        ipm = false;
                    }
                }
            }
            // first pass: a robot that has not converged after ipm_inline iterations leaves the loop here and is parked below
            // (while the store has room: a robot that finds it full simply carries on in this pass)
            if (ipm && mode == 0 && park && df.ipm_buf && it >= df.ipm_inline && *(volatile const int*)df.ipm_count < df.ipm_cap)
                ipm = false, park_ipm = true;
        }
        // parked for the interior-point pass: the step just computed is applied (as the next sweep A would have) and the new
        // interior point (u, s, lam) stored
        if (park_ipm) {
            int qq;
#ifdef __CUDA_ARCH__
            qq = atomicAdd(df.ipm_count, 1);
#else
            qq = (*df.ipm_count)++;
#endif
            if (qq < df.ipm_cap) {
                constexpr int PB = 5 + 2 * NR;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
                    double u[LB], x5[LB], d5[LB];
#pragma unroll
                    for (int c = 0; c < LB; ++c) u[c] = r[L::o_u + c], x5[c] = r[L::o_xv + c], d5[c] = r[L::o_du + c];
                    for_rows([&](auto tag, int slot, int) {
                        const double s = r[L::o_s + slot], lm = r[L::o_l + slot];
                        const RowStep qr = row_affine(rdot(tag, u), rdot(tag, x5), s, lm, rrhs(tag));
                        const double wc = (qr.dsa * qr.dla - tgt) * qr.is;
                        const double ds = -qr.rp - rdot(tag, d5), dl = -lm - wc - qr.d * ds;
                        parked_ipm(qq, v * PB + 5 + slot) = (float)(s + alpha * ds);
                        parked_ipm(qq, v * PB + 5 + NR + slot) = (float)(lm + alpha * dl);
                    });
#pragma unroll
                    for (int c = 0; c < LB; ++c) parked_ipm(qq, v * PB + c) = (float)(u[c] + alpha * d5[c]);
                }
                parked_ipm(qq, S * PB) = (float)gs;
                parked_ipm(qq, S * PB + 1) = (float)it;
                df.ipm_list[qq] = inst;  // (status stays 1 until a later pass certifies the robot)
            } else {
                status = 1;  // lost the race for the last places of the store: the warp-per-robot kernels take the robot
            }
        }

        // ---- 3. active-set polish + certificate (one attempt; anything else goes to the warp-per-robot kernels) ----
        bool pol = act && status == 0, polished = false, park_pol = false;
        int round = 0;
        if (mode == 1) {
            // polish pass: the parked state instead of an interior-point iterate
            pol = act;
            if (pol) {
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
                    r[L::o_am] = (double)parked(q, v * (NR + 1));
                    for_rows([&](auto, int slot, int) { r[L::o_l + slot] = (double)parked(q, v * (NR + 1) + 1 + slot); });
                }
                gs = (double)parked(q, S * (NR + 1));
                rdmax = (double)parked(q, S * (NR + 1) + 1);
                it = (int)parked(q, S * (NR + 1) + 2);
                status = 1;
            }
            round = df.inline_rounds;
        } else if (pol) {
            // diag(Hc) from the uncontrolled cost-to-go (into tv)
            init_P();
#pragma unroll 1
            for (int i = HZ - 1; i >= 0; --i) {
                SV P = Ps;
#pragma unroll
                for (int li = 0; li < NF; ++li) {
                    const int v = i * NF + li, l = foot_of(v);
                    SV r = rec(v);
                    double B3[3][LB];
                    load_B3(r, B3);
#pragma unroll
                    for (int c = 0; c < LB; ++c) {
                        double b6[6] = {B3[0][c], B3[1][c], B3[2][c], 0.0, 0.0, 0.0};
                        if (c < 3) b6[3 + c] = vm;
                        double acc = Rw(l, c);
#pragma unroll
                        for (int k = 0; k < 6; ++k)
#pragma unroll
                            for (int k2 = 0; k2 < 6; ++k2)
                                if ((k < 3 || k == 3 + c) && (k2 < 3 || k2 == 3 + c)) acc += b6[k] * P[pk(6 + k, 6 + k2)] * b6[k2];
                        r[L::o_tv + c] = acc;
                    }
                }
                if (i == 0) break;
                double r9[9];
                load_r9(rec(i * NF), dt, r9);
                congruence_pk(r9, dt);
#pragma unroll
                for (int a = 0; a < 12; ++a) P[pk(a, a)] += p.Q[a];
            }
            // hand-over: apply the last step to (s, lam); rows whose barrier weight dominates the curvature along them are active
#pragma unroll 1
            for (int v = 0; v < S; ++v) {
                SV r = rec(v);
                double u[LB], x5[LB], d5[LB], hd[LB];
#pragma unroll
                for (int c = 0; c < LB; ++c) u[c] = r[L::o_u + c], x5[c] = r[L::o_xv + c], d5[c] = r[L::o_du + c], hd[c] = r[L::o_tv + c];
                unsigned mk = 0u;
                for_rows([&](auto tag, int slot, int k) {
                    double s = r[L::o_s + slot], lm = r[L::o_l + slot];
                    const RowStep q = row_affine(rdot(tag, u), rdot(tag, x5), s, lm, rrhs(tag));
                    const double wc = (q.dsa * q.dla - tgt) * q.is;
                    const double ds = -q.rp - rdot(tag, d5), dl = -lm - wc - q.d * ds;
                    s += alpha * ds;
                    lm += alpha * dl;
                    r[L::o_l + slot] = lm;
                    double c5[LB];
                    rcoef(tag, c5);
                    double th = 0.0, aa = 0.0;
#pragma unroll
                    for (int c = 0; c < LB; ++c) th += c5[c] * c5[c] * hd[c], aa += c5[c] * c5[c];
                    if (lm * fmax(aa * aa, 1e-300) > th * s) mk |= 1u << k;
                });
                r[L::o_am] = (double)mk;
            }
            status = 1;
        }
#pragma unroll 1
        while (group_any_polish(pol)) {
            if (pol && round >= (HZ > 10 ? 3 : 1) * p.polish_rounds) pol = false;  // (long horizons: more weakly active rows to walk through)
            ++round;
            double pv[12];
            if (pol) {
                bool bad_blk = false;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
                    if (!blk_nullspace(r, (unsigned)r[L::o_am])) bad_blk = true;
                }
                if (bad_blk) pol = false;
#ifdef BMPC_LANE_DEBUG
                printf("[polish] round %d bad_blk %d\n", round, (int)bad_blk);
#endif
            }
            if (pol) {
                grad(L::o_pp, L::o_tv, nullptr);
                // reduced LQR: inputs w_b with u_b = p_b + N_b w_b; backward sweep (factor + backward half), then forward
                init_P();
#pragma unroll
                for (int a = 0; a < 12; ++a) pv[a] = 0.0;
            }
#pragma unroll 1
            for (int v = S - 1; v >= 0; --v) {
                stage_sync_polish(v == S - 1);
                if (!pol) continue;
                prefetch_rec(v - 1, L::REC);
                SV r = rec(v);
                const int l = foot_of(v);
                double N[LB * LB], G[15], rhs[LB];
                const int dim = (int)r[L::o_dim];
#pragma unroll
                for (int e = 0; e < LB * LB; ++e) N[e] = r[L::o_Nn + e];
#pragma unroll
                for (int c = 0; c < LB; ++c) rhs[c] = -r[L::o_tv + c];
                Nt_mul(N, dim, rhs);
#pragma unroll
                for (int a = 0; a < LB; ++a)
#pragma unroll
                    for (int b = 0; b <= a; ++b) G[tri(a, b)] = (a == b) ? Rw(l, a) : 0.0;
                if (!factor_stage<true>(v, r, G, rhs, pv, L::o_xv, N, dim)) {
                    pol = false;
#ifdef BMPC_LANE_DEBUG
                    printf("[polish] round %d factor breakdown at block %d (dim %d)\n", round, v, dim);
#endif
                }
            }
            bool changed = false;
            {
                double z[12];
#pragma unroll
                for (int a = 0; a < 12; ++a) z[a] = 0.0;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    stage_sync_polish(v == 0);
                    if (!pol) continue;
                    prefetch_rec(v + 1, L::REC);
                    SV r = rec(v);
                    double xs[LB], us[LB], up[LB];
                    solve_fwd_stage<true>(v, r, L::o_xv, z, xs);
#pragma unroll
                    for (int c = 0; c < LB; ++c) {
                        double acc = 0.0;
#pragma unroll
                        for (int a = 0; a < LB; ++a) acc += r[L::o_Nn + c * LB + a] * xs[a];
                        us[c] = acc;
                        up[c] = r[L::o_pp + c] + acc;
                        r[L::o_u + c] = up[c];
                    }
                    advance_z(v, r, us, z);
                    // primal check: violated inactive rows join the active set
                    unsigned mk = (unsigned)r[L::o_am];
                    bool ch = false;
                    for_rows([&](auto tag, int, int k) {
                        const double bk = rrhs(tag);
                        const double viol = rdot(tag, up) - bk;
                        if (viol > 1e-9 * (1.0 + fabs(bk)) && !((mk >> k) & 1u)) mk |= 1u << k, ch = true;
                    });
                    if (ch) r[L::o_am] = (double)mk, changed = true;
                }
#ifdef BMPC_LANE_DEBUG
                if (pol) printf("[polish] round %d primal: changed %d\n", round, (int)changed);
#endif
            }
            if (pol && !changed) {
                // dual check: minus the gradient must be a non-negative combination of the active rows
                grad(L::o_u, L::o_tv, nullptr);
                bool fail = false, released = false;
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
                    const unsigned mk = (unsigned)r[L::o_am];
                    unsigned drop = 0u;
                    if (blk_dual_fast(r, mk, gs)) continue;
                    if (!blk_dual_exact(r, mk, gs, &drop)) {
                        // releasing several rows of many blocks at once can cycle (release, re-add as violated, release ...):
                        // after the first rounds only one row per block is released at a time
                        if (round > 3 && drop != 0u) drop &= (~drop + 1u);
                        // ... and if that still cycles (several blocks exchanging the same pair of rows in step: seen on a degenerate
                        // standing instance at h = 30, period five rounds), from round 33 on ONE release per round in the whole
                        // problem, lowest block and lowest row first - the single exchange of a textbook active-set method (Bland's
                        // rule).  Not earlier: an instance that has to release rows in many blocks (the other degenerate h = 30
                        // instance: 31 rounds) would take a round per row (70 rounds with the rule from round 9 on).
                        if (drop == 0u) fail = true;
                        else if (round > 32 && released) changed = true;
                        else r[L::o_am] = (double)(mk & ~drop), changed = true, released = true;
#ifdef BMPC_LANE_DEBUG
                        printf("[polish] round %d block %d mask 0x%x dual check: drop 0x%x gs %.3e\n", round, v, mk, drop, gs);
#endif
                    }
                }
#ifdef BMPC_LANE_DEBUG
                printf("[polish] round %d dual: fail %d changed %d\n", round, (int)fail, (int)changed);
#endif
                if (fail) pol = false;
                else if (!changed) polished = true, pol = false;
            }
            // a robot that needs more rounds than this pass runs leaves the loop here and is parked below
            if (pol && mode != 1 && park && df.buf && round >= df.inline_rounds && *(volatile const int*)df.count < df.cap)
                pol = false, park_pol = true;
        }
        // parked for the polish pass: active-row masks and multipliers of every block
        if (park_pol) {
            int qq;
#ifdef __CUDA_ARCH__
            qq = atomicAdd(df.count, 1);
#else
            qq = (*df.count)++;
#endif
            if (qq < df.cap) {
#pragma unroll 1
                for (int v = 0; v < S; ++v) {
                    SV r = rec(v);
                    parked(qq, v * (NR + 1)) = (float)r[L::o_am];
                    for_rows([&](auto, int slot, int) { parked(qq, v * (NR + 1) + 1 + slot) = (float)r[L::o_l + slot]; });
                }
                parked(qq, S * (NR + 1)) = (float)gs;
                parked(qq, S * (NR + 1) + 1) = (float)rdmax;
                parked(qq, S * (NR + 1) + 2) = (float)it;
                df.list[qq] = inst;  // (status 1 until the polish pass certifies the robot)
            }                        // (store full: status 1, the warp-per-robot kernels take the robot)
        }
        if (act && !polished) {
            io.status[inst] = 1;
            io.iters[inst] = it;
            act = false;
        }
        if (!act) return;

        // ---- 4. outputs ----
        double u0[12];
        double umax = 0.0;
#pragma unroll 1
        for (int s = 0; s < HZ; ++s) {
            double u12[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) u12[e] = 0.0;
#pragma unroll
            for (int li = 0; li < NF; ++li) {
                const int v = s * NF + li, l = foot_of(v);
                SV r = rec(v);
                double up[LB];
#pragma unroll
                for (int c = 0; c < LB; ++c) up[c] = r[L::o_u + c];
#pragma unroll
                for (int c = 0; c < 3; ++c) {  // u = [f1 f2 m1 m2] (MPC.py:10)
                    if (l) u12[3 + c] = up[c]; else u12[c] = up[c];
                }
                if (l) u12[9] = pin, u12[10] = up[3], u12[11] = up[4];
                else u12[6] = pin, u12[7] = up[3], u12[8] = up[4];
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                io.controls[(size_t)inst * HZ * 12 + s * 12 + e] = u12[e];
                umax = fmax(umax, fabs(u12[e]));
                if (s == 0) u0[e] = u12[e];
            }
        }
        const double uscale = fmax(1.0, umax);
        if (io.states) grad(L::o_u, L::o_tv, io.states + (size_t)inst * HZ * 13);
        if (io.fric_active) {
            const double tol = 1e-6 * uscale;
#pragma unroll 1
            for (int s = 0; s < HZ; ++s) {
                unsigned mask = 0;
#pragma unroll
                for (int li = 0; li < NF; ++li) {
                    const int v = s * NF + li, l = foot_of(v);
                    SV r = rec(v);
                    const double f0 = r[L::o_u + 0], f1 = r[L::o_u + 1], f2 = r[L::o_u + 2];
                    if (f2 <= tol) continue;
                    const double res[4] = {f0 - p.mu * f2, f1 - p.mu * f2, -f0 - p.mu * f2, -f1 - p.mu * f2};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (res[q] >= -tol) mask |= 1u << (4 * l + q);
                }
                io.fric_active[(size_t)inst * HZ + s] = (uint8_t)mask;
            }
        }
        if (io.do_lowlevel && io.tau) {
            double q[10], qd[10], pf[6], x_fb[12], rotn[9];
            for (int a = 0; a < 12; ++a) x_fb[a] = xfb[a];
            eul2rotm(x_fb, rotn);
            for (int a = 0; a < 10; ++a) q[a] = io.q[(size_t)inst * 10 + a], qd[a] = io.qd[(size_t)inst * 10 + a];
            for (int a = 0; a < 6; ++a) pf[a] = io.pf_w[(size_t)inst * 6 + a];
            for (int leg = 0; leg < 2; ++leg) {
                double tl[5];
                lowlevel_leg_inl(p, x_fb, io.t_swing[inst], pf, q, qd, rotn, leg, (double)cont0[leg], u0, tl);
                for (int c = 0; c < 5; ++c) io.tau[(size_t)inst * 10 + 5 * leg + c] = tl[c];
            }
        }
        if (io.ws_mask) {
#pragma unroll 1
            for (int s = 0; s < HZ; ++s) {
                int m2[2] = {-1, -1};
#pragma unroll
                for (int li = 0; li < NF; ++li) {
                    const int v = s * NF + li;
                    const int mk = (int)rec(v)[L::o_am];
                    if (foot_of(v)) m2[1] = mk; else m2[0] = mk;
                }
                io.ws_mask[(size_t)inst * 2 * HZ + 2 * s] = m2[0];
                io.ws_mask[(size_t)inst * 2 * HZ + 2 * s + 1] = m2[1];
            }
        }
        io.status[inst] = 0;
        io.iters[inst] = it;
        if (io.resid) io.resid[2 * inst] = 0.0, io.resid[2 * inst + 1] = rdmax;
    }
};

#if defined(__CUDACC__) && !defined(BMPC_LANE_HOST_ONLY)
// Work distribution: a global counter hands out slices of the class's work list - 32 robots per warp when the warps run
// independently (p.lane_sync == 0), 32 x (warps per CTA) robots per CTA when they run in lockstep.
// Dynamic shared memory: per warp the packed cost-to-go of its 32 robots (78 x 32 doubles).
template <int HZ, int NF, unsigned RM>
__global__ void __launch_bounds__(256, 1) lane_tick_kernel(const __grid_constant__ DevParams p, const IoPtrs io,
                                                           const int* __restrict__ work_list, const int* __restrict__ work_count,
                                                           int* __restrict__ slice_counter, double* __restrict__ wsbase, int min_count,
                                                           const __grid_constant__ LaneDefer df, int mode) {
    using L = LaneRec<HZ, NF>;
    extern __shared__ double lane_smem[];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int warp = blockIdx.x * nw + wib;
    // modes 1 and 2 (later passes): the work list is the list of the robots parked for that pass
    const int count = mode == 1 ? min(*df.count, df.cap) : mode == 2 ? min(*df.ipm_count, df.ipm_cap) : *work_count;
    if (mode == 1) work_list = df.list;
    if (mode == 2) work_list = df.ipm_list;
    // a class too small to be worth 32-robot slices stays on the warp-per-robot kernel: collect_or_all_kernel hands the whole list over
    if (mode == 0 && count < min_count) return;
    SV ws{wsbase + (size_t)warp * L::total * 32 + lane};
    SV ps{lane_smem + (size_t)wib * L::smem_doubles * 32 + lane};
    while (true) {
        int base = 0;
        if (p.lane_sync & 3) {
            __syncthreads();  // (everybody has read s_base of the previous round)
            if (threadIdx.x == 0) s_base = atomicAdd(slice_counter, 32 * nw);
            __syncthreads();
            base = s_base + 32 * wib;
            if (s_base >= count) break;
        } else {
            if (lane == 0) base = atomicAdd(slice_counter, 32);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= count) break;
        }
        LaneSolver<HZ, NF, true, RM> solver(p, ws, ps, lane, df);
        solver.mode = mode;
        solver.park = mode == 2 || (mode == 0 && count >= df.min_count);
        solver.run(io, base + lane < count ? work_list[base + lane] : -1, base + lane);
        __syncwarp();
    }
}
#endif

}  // namespace bmpc
