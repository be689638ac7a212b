// The fused per-instance MPC tick kernel: assembly + interior point + active-set polish +
// torque map.  One thread GROUP (a warp for the walking class, a 128-thread CTA for the
// standing class) owns one robot at a time and strides over its class's work list.
//
// Data layout in shared memory.  The contact-reduced condensed Hessian Hc (n = LB*S, S stance
// foot-stages) is stored as the lower BLOCK triangle of LB x LB tiles, TS doubles apart
// (LB*LB rounded up to an even count so every tile is 16-byte aligned and a lane's tile is
// 13 LDS.128; tile stride 208 B is conflict free across a quarter warp).  All dense work is
// tile-granular with the tile in registers: a lane owns whole tiles, loops are fully unrolled,
// there is no per-element index arithmetic.
//
// H itself is kept once per resident group in a global scratch (L2 resident: 148 SMs x ~10
// groups x 11 KB); the factor is built in place in shared memory and H is brought back by ONE
// TMA bulk copy per iteration (cp.async.bulk + mbarrier), overlapped with the step-length work.
#pragma once
#include "bmpc_kernels.cuh"
#include "bmpc_polish.cuh"
#include "bmpc_riccati.cuh"

namespace bmpc {

// ------------------------------------------------------------------------------------
// group primitives (NT == 32: warp-synchronous, no CTA barrier anywhere)
// ------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void gsync() {
    if constexpr (NT == 32) __syncwarp();
    else __syncthreads();
}
template <int NT>
__device__ __forceinline__ double gsum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if constexpr (NT == 32) return v;
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    __syncthreads();
    return t;
}
template <int NT>
__device__ __forceinline__ double gmax(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if constexpr (NT == 32) return v;
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) t = fmax(t, red[w]);
    __syncthreads();
    return t;
}
template <int NT>
__device__ __forceinline__ float gmaxf(float v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if constexpr (NT == 32) return v;
    float* redf = reinterpret_cast<float*>(red);
    if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = redf[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) t = fmaxf(t, redf[w]);
    __syncthreads();
    return t;
}
// -dv / v as a float, for the step-length ratio test (a relative error of 1e-7 in a step length that is
// cut by 0.995 afterwards is harmless).  NaN/Inf propagate so a broken direction is still detected.
__device__ __forceinline__ float step_ratio(double dv, double v) {
    return __fdividef(-(float)dv, (float)v);
}
template <int NT>
__device__ __forceinline__ int gany(int pred) {
    if constexpr (NT == 32) return __any_sync(0xffffffffu, pred);
    else return __syncthreads_or(pred);
}

// ------------------------------------------------------------------------------------
// tiles
// ------------------------------------------------------------------------------------
template <int LB>
struct TileT {
    static constexpr int E = LB * LB;
    static constexpr int TS = (E + 2) & ~1;  // 26 doubles (LB=5), 38 (LB=6): 16-byte aligned, bank-spread
};
__device__ __forceinline__ int tidx(int jr, int jc) { return (jr * (jr + 1)) / 2 + jc; }  // jc <= jr

template <int LB>
__device__ __forceinline__ void tile_load(const double* __restrict__ p, double (&t)[LB * LB]) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < (LB * LB) / 2; ++i) {
        const double2 v = q[i];
        t[2 * i] = v.x;
        t[2 * i + 1] = v.y;
    }
    if constexpr ((LB * LB) & 1) t[LB * LB - 1] = p[LB * LB - 1];
}
template <int LB>
__device__ __forceinline__ void tile_store(double* __restrict__ p, const double (&t)[LB * LB]) {
    double2* q = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int i = 0; i < (LB * LB) / 2; ++i) q[i] = make_double2(t[2 * i], t[2 * i + 1]);
    if constexpr ((LB * LB) & 1) p[LB * LB - 1] = t[LB * LB - 1];
}

// Cholesky of a symmetric LB x LB tile (lower part used).  Returns the factor in l (lower part)
// with the RECIPROCALS of its diagonal on the diagonal (that is what every consumer needs).
// Every lane of the group runs this on the same data (a broadcast read), so the result is in
// registers everywhere and needs no exchange.  false on a non-positive pivot.
template <int LB>
__device__ __forceinline__ bool tile_chol(const double (&a)[LB * LB], double (&l)[LB * LB]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < LB; ++j) {
        double djj = a[j * LB + j];
#pragma unroll
        for (int k = 0; k < j; ++k) djj -= l[j * LB + k] * l[j * LB + k];
        ok = ok && (djj > 0.0) && (djj < 1e300);
        const double r = rsqrt(djj);
        l[j * LB + j] = r;
#pragma unroll
        for (int i = j + 1; i < LB; ++i) {
            double v = a[i * LB + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= l[i * LB + k] * l[j * LB + k];
            l[i * LB + j] = v * r;
        }
#pragma unroll
        for (int i = 0; i < j; ++i) l[i * LB + j] = 0.0;
    }
    return ok;
}

// In-place block Cholesky of the tile matrix in Mb (S block rows), left in the SOLVE form
// M = Lh D Lh' with Lh unit block-lower: tile (jr,jc), jr > jc, holds Lh(jr,jc) = L(jr,jc) inv(L(jc,jc))
// and the diagonal tile (j,j) holds the Cholesky factor of D_j with reciprocal diagonal.
// Uniform return value.
template <int LB, int NT>
__device__ __noinline__ bool tile_factor(double* __restrict__ Mb, int S) {
    constexpr int E = LB * LB, TS = TileT<LB>::TS;
    const int tid = threadIdx.x & (NT - 1);
    bool ok = true;
    for (int jc = 0; jc < S; ++jc) {
        double l[E];
        {
            double dg[E];
            tile_load<LB>(Mb + tidx(jc, jc) * TS, dg);
            ok = tile_chol<LB>(dg, l) && ok;
        }
        gsync<NT>();  // everyone has read the diagonal tile before it is overwritten
        if (!ok) return false;  // uniform: every lane factored the same tile
        const int below = S - 1 - jc;
        // panel: L(jr,jc) = A(jr,jc) inv(L(jc,jc))'  (forward substitution along each row)
        for (int t = tid; t < below; t += NT) {
            double* tp = Mb + tidx(jc + 1 + t, jc) * TS;
            double a[E];
            tile_load<LB>(tp, a);
#pragma unroll
            for (int c = 0; c < LB; ++c)
#pragma unroll
                for (int r = 0; r < LB; ++r) {
                    double v = a[r * LB + c];
#pragma unroll
                    for (int k = 0; k < c; ++k) v -= a[r * LB + k] * l[c * LB + k];
                    a[r * LB + c] = v * l[c * LB + c];
                }
            tile_store<LB>(tp, a);
        }
        if (tid == NT - 1) tile_store<LB>(Mb + tidx(jc, jc) * TS, l);
        gsync<NT>();
        // trailing update: A(jr,jc2) -= L(jr,jc) L(jc2,jc)'   for jc < jc2 <= jr
        const int ntr = below * (below + 1) / 2;
        for (int t = tid; t < ntr; t += NT) {
            int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            if (r * (r + 1) / 2 > t) --r;
            if ((r + 1) * (r + 2) / 2 <= t) ++r;
            const int c = t - r * (r + 1) / 2;
            const int jr = jc + 1 + r, jc2 = jc + 1 + c;
            double* tp = Mb + tidx(jr, jc2) * TS;
            double acc[E];
            tile_load<LB>(tp, acc);
            const double* ap = Mb + tidx(jr, jc) * TS;
            const double* bp = Mb + tidx(jc2, jc) * TS;
            // rolled over k: the body stays resident in the instruction cache (the kernel is fetch-bound)
#pragma unroll 1
            for (int k = 0; k < LB; ++k) {
                double a[LB], b[LB];
#pragma unroll
                for (int x = 0; x < LB; ++x) a[x] = ap[x * LB + k], b[x] = bp[x * LB + k];
#pragma unroll
                for (int x = 0; x < LB; ++x)
#pragma unroll
                    for (int y = 0; y < LB; ++y) acc[x * LB + y] -= a[x] * b[y];
            }
            tile_store<LB>(tp, acc);
        }
        gsync<NT>();
    }
    // solve form: Lh(jr,jc) = L(jr,jc) inv(L(jc,jc))  (backward substitution along each row)
    const int nlow = S * (S - 1) / 2;
    for (int t = tid; t < nlow; t += NT) {
        int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        if (r * (r + 1) / 2 > t) --r;
        if ((r + 1) * (r + 2) / 2 <= t) ++r;
        const int jc = t - r * (r + 1) / 2, jr = r + 1;  // strictly lower tiles
        double* tp = Mb + tidx(jr, jc) * TS;
        double a[E], l[E];
        tile_load<LB>(tp, a);
        tile_load<LB>(Mb + tidx(jc, jc) * TS, l);
#pragma unroll
        for (int c = LB - 1; c >= 0; --c)
#pragma unroll
            for (int x = 0; x < LB; ++x) {
                double v = a[x * LB + c];
#pragma unroll
                for (int k = c + 1; k < LB; ++k) v -= a[x * LB + k] * l[k * LB + c];
                a[x * LB + c] = v * l[c * LB + c];
            }
        tile_store<LB>(tp, a);
    }
    gsync<NT>();
    return ok;
}

// x <- inv(M) x with the factor of tile_factor (unit block-lower Lh below the diagonal, chol(D_j)
// on it).  x has LB*S entries in shared memory.  One group barrier per block step.
template <int LB, int NT, int NPT>
__device__ __noinline__ void tile_solve(const double* __restrict__ Mb, int S, double* __restrict__ x) {
    constexpr int TS = TileT<LB>::TS;
    const int tid = threadIdx.x & (NT - 1);
    const int n = S * LB;
    // forward: z_jr -= Lh(jr,jc) z_jc for jr > jc
    for (int jc = 0; jc + 1 < S; ++jc) {
        const double* zc = x + jc * LB;
        const int cnt = (S - 1 - jc) * LB;
#pragma unroll 1
        for (int t = tid; t < cnt; t += NT) {
            const int r = t / LB, a = t - r * LB;
            const double* lt = Mb + tidx(jc + 1 + r, jc) * TS + a * LB;
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < LB; ++b) acc += lt[b] * zc[b];
            x[(jc + 1 + r) * LB + a] -= acc;
        }
        gsync<NT>();
    }
    // middle: w_j = inv(D_j) z_j by the two triangular substitutions with chol(D_j); one lane per
    // block, which reads and writes only its own LB entries (no hazard, one barrier)
    for (int j = tid; j < S; j += NT) {
        double l[LB * LB], z[LB];
        tile_load<LB>(Mb + tidx(j, j) * TS, l);
#pragma unroll
        for (int a = 0; a < LB; ++a) z[a] = x[j * LB + a];
#pragma unroll
        for (int a = 0; a < LB; ++a) {
#pragma unroll
            for (int b = 0; b < a; ++b) z[a] -= l[a * LB + b] * z[b];
            z[a] *= l[a * LB + a];
        }
#pragma unroll
        for (int a = LB - 1; a >= 0; --a) {
#pragma unroll
            for (int b = a + 1; b < LB; ++b) z[a] -= l[b * LB + a] * z[b];
            z[a] *= l[a * LB + a];
        }
#pragma unroll
        for (int a = 0; a < LB; ++a) x[j * LB + a] = z[a];
    }
    gsync<NT>();
    // backward: x_r -= Lh(jc,r)' x_jc for r < jc
    for (int jc = S - 1; jc > 0; --jc) {
        const double* xc = x + jc * LB;
        const int cnt = jc * LB;
#pragma unroll 1
        for (int t = tid; t < cnt; t += NT) {
            const int r = t / LB, a = t - r * LB;
            const double* lt = Mb + tidx(jc, r) * TS + a;
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < LB; ++b) acc += lt[b * LB] * xc[b];
            x[r * LB + a] -= acc;
        }
        gsync<NT>();
    }
}

// out = H v (+ add) from the symmetric tile matrix (diagonal tiles stored in full)
template <int LB, int NT>
__device__ __noinline__ void tile_symv(const double* __restrict__ Mb, int S, const double* __restrict__ v,
                          const double* __restrict__ add, double* __restrict__ out) {
    constexpr int TS = TileT<LB>::TS;
    const int n = S * LB;
    for (int i = threadIdx.x & (NT - 1); i < n; i += NT) {
        const int j = i / LB, a = i - j * LB;
        double acc = add ? add[i] : 0.0;
#pragma unroll 1
        for (int jc = 0; jc <= j; ++jc) {
            const double* tp = Mb + tidx(j, jc) * TS + a * LB;
            const double* vv = v + jc * LB;
#pragma unroll
            for (int b = 0; b < LB; ++b) acc += tp[b] * vv[b];
        }
#pragma unroll 1
        for (int jr = j + 1; jr < S; ++jr) {
            const double* tp = Mb + tidx(jr, j) * TS + a;
            const double* vv = v + jr * LB;
#pragma unroll
            for (int b = 0; b < LB; ++b) acc += tp[b * LB] * vv[b];
        }
        out[i] = acc;
    }
}

// ------------------------------------------------------------------------------------
// shared-memory layout (offsets in doubles)
// ------------------------------------------------------------------------------------
// MG: the tile matrix does not fit in shared memory (h = 30 standing: 1,830 tiles, 380 KB) and lives in
// the per-group global scratch next to H (L2 resident); everything else is unchanged.
// RIC: stage-wise (Riccati) backend (bmpc_riccati.cuh): no tile matrix at all; the o_M region holds the per-block
// feedback rows, inverse input Hessians and input maps instead.
template <int HZ, int SMAX, int LB, bool MG = false, bool RIC = false>
struct TickLayout {
    static constexpr int TS = TileT<LB>::TS;
    static constexpr int N = LB * SMAX;
    static constexpr int NTILE = SMAX * (SMAX + 1) / 2;
    static constexpr int MB = NTILE * TS;              // tile matrix (H, then the factor, in place)
    static constexpr int NV = (N + 3) & ~3;
    static constexpr int NPAIR = HZ * (HZ + 1) / 2;
    static constexpr int NAB = LB * (LB + 1) / 2;
    // ---- shared memory ----
    static constexpr int o_M = 0;
    static constexpr int RB = (Ric<LB, HZ, SMAX>::total + 1) & ~1;
    static constexpr int o_in = o_M + (RIC ? RB : (MG ? 0 : MB));   // 2 x IN_DOUBLES TMA destinations (inputs, double buffered)
    static constexpr int o_g = o_in + 2 * IN_DOUBLES;
    static constexpr int o_u = o_g + NV;
    static constexpr int o_Cb = o_u + NV;
    static constexpr int o_rb = o_Cb + MAXROWS * LB;
    static constexpr int o_ub = o_rb + MAXROWS + 2;
    static constexpr int o_red = o_ub + 8;
    static constexpr int o_int = o_red + 16;
    static constexpr int n_int = 2 * SMAX + 2 * HZ + HZ + 2 * HZ + 2 * SMAX + 16 + 2 * HZ;
    static constexpr int o_bar = ((o_int + (n_int + 1) / 2 + 1) + 1) & ~1;  // mbarriers, 16-byte aligned
    // work region: three n-vectors + 6 row arrays (s, lam, lam/s, rp|ds, wc|dlam, w; run-time sized).
    // The assembly phase, which runs before any of them is live, uses the same region for the
    // per-stage dynamics pieces (a_*); the polish reuses the four dead row arrays (Nn, tvp).
    static constexpr int o_work = o_bar + 4;
    static constexpr int o_du = o_work;                // step / polish: particular solution
    static constexpr int o_x = o_du + NV;              // H u + g, residual, right-hand side, solution
    static constexpr int o_up = o_x + NV;              // polished point
    static constexpr int o_rd = o_up + NV;             // stationarity residual (carried by its recurrence)
    static constexpr int o_rows = o_rd + NV;
    static constexpr int a_xref = o_work;
    static constexpr int a_rinv = a_xref + HZ * 12;
    static constexpr int a_psum = a_rinv + HZ * 9;
    static constexpr int a_iwinv = a_psum + HZ * 9;
    static constexpr int a_err = a_iwinv + HZ * 9;
    static constexpr int a_footv = a_err + HZ * 12;
    static constexpr int a_rot = a_footv + 24;
    static constexpr int a_W = a_rot + 12;
    static constexpr int a_Wp = a_W + SMAX * 3 * LB;
    static constexpr int a_Vp = a_Wp + SMAX * 3;
    static constexpr int a_end = a_Vp + SMAX * 3;
    // rows per array for `mb` inequality rows per block (even, so every array stays 16-byte aligned)
    __host__ __device__ static constexpr int row_stride(int mb) { return (mb * SMAX + 1) & ~1; }
    __host__ __device__ static constexpr int imax(int a, int b) { return a > b ? a : b; }
    __host__ __device__ static constexpr int rows_doubles(int mb) {
        return imax(imax(6 * row_stride(mb), 2 * row_stride(mb) + SMAX * LB * LB + NV), a_end - o_rows);
    }
    __host__ __device__ static constexpr size_t bytes(int mb) { return size_t(o_rows + rows_doubles(mb)) * 8; }
    // ---- per-group global scratch (L2 resident): H, the stage-pair kernels, and what only the output
    //      stage needs again (saved after assembly, read back once) ----
    static constexpr int g_H = 0;
    static constexpr int g_M = g_H + MB;               // MG only: the working copy (H + barrier terms, then the factor)
    static constexpr int g_pairs = g_M + (MG ? MB : 0);  // stage-pair kernels T (9) + cpp (1)
    static constexpr int g_save = g_pairs + NPAIR * 10;  // copy of the a_* region
    static constexpr int g_total = (g_save + (a_end - o_work) + 1) & ~1;
};

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
// NW > 1 (walking class, NT == 32): NW independent robots per CTA, one warp each, kept in loose
// lockstep by one mbarrier arrival per iteration so that the warps of an SM run the same code at
// about the same time and share instruction-cache lines (the kernel is instruction-fetch bound
// otherwise: profiles/r1_summary.md).
template <int HZ, int SMAX, int LB, int NT, int NW, bool MG = false, bool RIC = false>
__global__ void __launch_bounds__(NT * NW) mpc_tick2_kernel(const __grid_constant__ DevParams p, const IoPtrs io,
                                                       const int* __restrict__ work_list,
                                                       const int* __restrict__ work_count,
                                                       double* __restrict__ hscratch) {
    using L = TickLayout<HZ, SMAX, LB, MG, RIC>;
    constexpr int E = LB * LB, TS = L::TS, NAB = L::NAB;
    constexpr int NPT = (LB * SMAX + NT - 1) / NT;  // variables per thread
    static_assert(NW == 1 || NT == 32, "several robots per CTA only with one warp per robot");
    extern __shared__ __align__(16) double sm_cta[];
    const int tid = threadIdx.x & (NT - 1);
    const int wid = threadIdx.x / NT;                                   // robot slot inside the CTA
    const int group = blockIdx.x * NW + wid, ngroups = gridDim.x * NW;  // persistent group id / count
    const int per_group = (int)(L::bytes(p.mb) / 8);
    double* sm = sm_cta + (size_t)wid * per_group;
    uint64_t* lockbar = reinterpret_cast<uint64_t*>(sm_cta + (size_t)NW * per_group);  // CTA-wide lockstep barrier

    // per-group scratch in global memory: plain (coherent) loads/stores, ordered by the group barrier
    double* gbase = hscratch + (size_t)group * L::g_total;
    double* Mb = MG ? gbase + L::g_M : sm + L::o_M;
    double* s_in = sm + L::o_in;
    double* gv = sm + L::o_g;
    double* uv = sm + L::o_u;
    double* duv = sm + L::o_du;
    double* xv = sm + L::o_x;
    double* upv = sm + L::o_up;
    double* rdv = sm + L::o_rd;
    double* ppv = duv;               // polish: particular solution (the step vector is dead then)
    const int mrs = L::row_stride(p.mb);
    double* r_s = sm + L::o_rows;
    double* r_l = r_s + mrs;
    double* r_d = r_l + mrs;
    double* r_p = r_d + mrs;
    double* r_c = r_p + mrs;
    double* r_w = r_c + mrs;
    double* Nn = r_d;                // polish: null-space blocks and one n-vector over the dead row arrays
    double* tvp = Nn + SMAX * E;
    double* Cb = sm + L::o_Cb;
    double* rb = sm + L::o_rb;
    double* ub = sm + L::o_ub;
    double* red = sm + L::o_red;
    int* blk_stage = reinterpret_cast<int*>(sm + L::o_int);
    int* blk_foot = blk_stage + SMAX;
    int* blockOf = blk_foot + SMAX;   // [HZ][2]
    int* footsel = blockOf + 2 * HZ;  // [HZ]
    int* cont = footsel + HZ;         // [HZ][2]
    int* amask = cont + 2 * HZ;       // [SMAX] active-row bit masks (polish)
    int* bdim = amask + SMAX;         // [SMAX] null-space dimension per block
    int* misc = bdim + SMAX;          // [0]=S  [1]=flag
    int* sfirst = misc + 16;          // [HZ] first block of each stage (Riccati backend)
    int* scnt = sfirst + HZ;          // [HZ] blocks per stage
    Ric<LB, HZ, SMAX> ric;  // stage-wise backend: one base pointer + compile-time offsets, passed by value
    ric.b = sm + L::o_M;
    ric.sfirst = sfirst, ric.scnt = scnt, ric.blk_foot = blk_foot;
    ric.dt = p.dt, ric.vm = p.dt / p.mass;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::o_bar);  // [0],[1] inputs, [2] H reload

    const int count = *work_count;
    uint32_t lockpar = 0u;
    // loose lockstep: every warp arrives once per "step" (each iteration, each polish round, output + the next
    // robot's assembly); a warp that runs out of work drops out of the barrier for good.  Steps of similar length
    // keep the waiting low: measured 49.0 ms (this) vs 51.0 ms (extra step at instance start) vs 52.5 ms (no lockstep)
    auto lock_sync = [&]() {
        if constexpr (NW > 1) {
            if (p.lock_mode == 2) return;
            __syncwarp();
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(lockbar)) : "memory");
            mbar_wait(lockbar, lockpar);
            lockpar ^= 1u;
        }
    };
    auto lock_drop = [&]() {
        if constexpr (NW > 1) {
            if (p.lock_mode == 2) return;
            __syncwarp();
            if (tid == 0) asm volatile("mbarrier.arrive_drop.shared::cta.b64 _, [%0];" ::"r"(smem_u32(lockbar)) : "memory");
        }
    };
    double* hglob = gbase + L::g_H;
    double* pairs = gbase + L::g_pairs;
    double* gsave = gbase + L::g_save;
    // assembly-phase arrays, aliased onto the work region
    double* xref = sm + L::a_xref;
    double* rinv = sm + L::a_rinv;
    double* psum = sm + L::a_psum;
    double* iwinv = sm + L::a_iwinv;
    double* err = sm + L::a_err;
    double* footv = sm + L::a_footv;
    double* rotn = sm + L::a_rot;
    double* Wm = sm + L::a_W;
    double* Wp = sm + L::a_Wp;
    double* Vp = sm + L::a_Vp;

    const int mb = p.mb;
    const double dt = p.dt;

    auto issue_loads = [&](int inst, int buf) {
        double* dst = s_in + buf * IN_DOUBLES;
        const uint32_t bytes = 96 + 48 + (io.do_lowlevel ? (80 + 80 + 48) : 0);
        mbar_expect_tx(&bars[buf], bytes);
        tma_load_1d(dst, io.x_fb + (size_t)inst * 12, 96, &bars[buf]);
        tma_load_1d(dst + 12, io.foot + (size_t)inst * 6, 48, &bars[buf]);
        if (io.do_lowlevel) {
            tma_load_1d(dst + 18, io.q + (size_t)inst * 10, 80, &bars[buf]);
            tma_load_1d(dst + 28, io.qd + (size_t)inst * 10, 80, &bars[buf]);
            tma_load_1d(dst + 38, io.pf_w + (size_t)inst * 6, 48, &bars[buf]);
        }
    };
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        if (NW > 1 && wid == 0) mbar_init(lockbar, NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (group >= count) {
        lock_drop();
        return;
    }
    if (io.use_tma && tid == 0) issue_loads(work_list[group], 0);
    uint32_t parity[2] = {0u, 0u};
    uint32_t hparity = 0u;
    int buf = 0;

    // dynamic work distribution: the first `ngroups` items are taken statically, every further one from a global counter
    // (robots take 7..17 iterations: static striding leaves the slowest group ~4 % behind the average)
    int* dyn_counter = const_cast<int*>(work_count) + 3;
    auto fetch_next = [&]() -> int {
        int v = 0;
        if constexpr (NT == 32) {
            if (tid == 0) v = ngroups + atomicAdd(dyn_counter, 1);
            v = __shfl_sync(0xffffffffu, v, 0);
        } else {
            if (tid == 0) misc[2] = ngroups + atomicAdd(dyn_counter, 1);
            __syncthreads();
            v = misc[2];
            __syncthreads();
        }
        return v;
    };
    for (int w = group, wn = 0; w < count; w = wn, buf ^= 1) {
        const int inst = work_list[w];
        wn = fetch_next();
        if (p.lock_mode == 0) lock_sync();  // default (mode 3): output of one robot and assembly of the next form ONE step
        // ---- 0. inputs ---------------------------------------------------------------
        double* cur = s_in + buf * IN_DOUBLES;
        if (io.use_tma) {
            mbar_wait(&bars[buf], parity[buf]);
            parity[buf] ^= 1u;
        } else {
            for (int i = tid; i < 44; i += NT) {
                double v = 0.0;
                if (i < 12) v = io.x_fb[(size_t)inst * 12 + i];
                else if (i < 18) v = io.foot[(size_t)inst * 6 + i - 12];
                else if (io.do_lowlevel) {
                    if (i < 28) v = io.q[(size_t)inst * 10 + i - 18];
                    else if (i < 38) v = io.qd[(size_t)inst * 10 + i - 28];
                    else v = io.pf_w[(size_t)inst * 6 + i - 38];
                }
                cur[i] = v;
            }
        }
        for (int i = tid; i < 2 * HZ; i += NT) cont[i] = io.contact[(size_t)inst * 2 * HZ + i] ? 1 : 0;
        gsync<NT>();
        if (io.use_tma && tid == 0 && wn < count) issue_loads(work_list[wn], buf ^ 1);

        const double* x_fb = cur;
        const double* foot = cur + 12;
        const int phase_k = io.phase_k[inst];

        // ---- 1. per-instance scalars, references, per-stage dynamics (MPC.py:61-109, 148-185)
        bool bad = false;
        for (int i = tid; i < 18; i += NT) bad = bad || !isfinite(cur[i]);
        for (int task = tid; task < 4 + HZ; task += NT) {
            if (task == 0) {
                int S = 0;  // block list: stance foot-stages in (stage, foot) order
                for (int s = 0; s < HZ; ++s) {
                    sfirst[s] = S < SMAX ? S : 0;
                    scnt[s] = (cont[2 * s] ? 1 : 0) + (cont[2 * s + 1] ? 1 : 0);
                    for (int l = 0; l < 2; ++l) {
                        int b = -1;
                        if (cont[2 * s + l]) {
                            if (S < SMAX) blk_stage[S] = s, blk_foot[S] = l, b = S;
                            ++S;
                        }
                        blockOf[2 * s + l] = b;
                    }
                }
                misc[0] = S;
            } else if (task == 1) {
                // next footholds (MPC.py:73-93), including the x_fb[10] quirk of MPC.py:87
                const double hh = (double)p.h;
                const double ex = p.kv * (x_fb[3] - p.x_cmd[3]), ey = p.kv * (x_fb[4] - p.x_cmd[4]);
                const double x1 = x_fb[3] + x_fb[9] * 1 / 2 * hh / 2 * dt + ex;
                const double x2 = x_fb[3] + x_fb[9] * 1 / 2 * hh * dt + ex;
                const double y1 = x_fb[4] + x_fb[10] * 1 / 2 * hh / 2 * dt + ey;
                const double y2 = x_fb[10] + x_fb[10] * 1 / 2 * hh * dt + ey;
                for (int c = 0; c < 6; ++c) footv[c] = foot[c];
                footv[6] = x1, footv[7] = y1, footv[8] = 0.0, footv[9] = x1, footv[10] = y1, footv[11] = 0.0;
                footv[12] = x2, footv[13] = y2, footv[14] = 0.0, footv[15] = x2, footv[16] = y2, footv[17] = 0.0;
            } else if (task == 2) {
                eul2rotm(x_fb, rotn);
            } else if (task == 3) {
                // strictly feasible start inside one block's polytope
                double u6[6];
                for (int c = 0; c < 6; ++c) {
                    const double lo = p.lo6[c], hi = p.hi6[c];
                    const double v0 = fmin(fmax(0.0, lo + 0.1 * (hi - lo)), hi - 0.1 * (hi - lo));
                    u6[c] = (hi > lo) ? v0 : lo;
                }
                u6[2] = p.lo6[2] + p.init_fz_frac * (p.hi6[2] - p.lo6[2]);
                for (int c = 0; c < 2; ++c) {
                    const double lo = fmax(p.lo6[c], -p.mu * u6[2]), hi = fmin(p.hi6[c], p.mu * u6[2]);
                    if (p.hi6[c] > p.lo6[c]) u6[c] = 0.5 * (lo + hi);
                }
                for (int c = 0; c < LB; ++c) ub[c] = u6[p.comps[c]];
            } else {
                const int k = task - 4;
                // foot reference variant per stage (MPC.py:99-108)
                const int kk = phase_k % 5;
                int sel = 0;
                if (cont[0] + cont[1] == 1) sel = (k < 5 - kk) ? 0 : ((k < 10 - kk) ? 1 : 2);
                footsel[k] = sel;
                // state reference column k (MPC.py:61-70)
                double xr[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) xr[i] = (k == 0) ? x_fb[i] : p.x_cmd[i];
                if (k > 0) {
#pragma unroll
                    for (int i = 0; i < 6; ++i)
                        if (p.x_cmd[i + 6] != 0.0) xr[i] = x_fb[i] + p.x_cmd[i + 6] * (k * dt);
                }
#pragma unroll
                for (int i = 0; i < 12; ++i) xref[12 * k + i] = xr[i];
                // dynamics read x[0] as yaw, x[1] pitch, x[2] roll (MPC.py:151-153)
                double sz, cz, sy, cy, sx, cx;
                sincos(xr[0], &sz, &cz);
                sincos(xr[1], &sy, &cy);
                sincos(xr[2], &sx, &cx);
                // Rot = Rx(roll) Ry(pitch) Rz(yaw)  (extrinsic 'zyx', MPC.py:156)
                double rot[9];
                rot[0] = cy * cz;
                rot[1] = -cy * sz;
                rot[2] = sy;
                rot[3] = sx * sy * cz + cx * sz;
                rot[4] = -sx * sy * sz + cx * cz;
                rot[5] = -sx * cy;
                rot[6] = -cx * sy * cz + sx * sz;
                rot[7] = cx * sy * sz + sx * cz;
                rot[8] = cx * cy;
                double tmp[9], iw[9], iwi[9];
                mat3_mul(p.inertia, rot, tmp);
                mat3_tmul(rot, tmp, iw);  // Rot' I Rot  (MPC.py:157)
                if (!mat3_inv(iw, iwi)) bad = true;
#pragma unroll
                for (int i = 0; i < 9; ++i) iwinv[9 * k + i] = iwi[i];
                // inverse of the euler-rate matrix at MPC.py:160-164, closed form
                const double icp = 1.0 / cy;  // pitch = +-pi/2 -> singular
                if (!isfinite(icp) || fabs(cy) < 1e-9) bad = true;
                double* ri = rinv + 9 * k;
                ri[0] = cz * icp, ri[1] = sz * icp, ri[2] = 0.0;
                ri[3] = -sz, ri[4] = cz, ri[5] = 0.0;
                ri[6] = cz * sy * icp, ri[7] = sz * sy * icp, ri[8] = 1.0;
            }
        }
        const int any_bad = gany<NT>(bad ? 1 : 0);
        gsync<NT>();
        const int S = misc[0];
        const int n = S * LB;
        const int m = S * mb;
        if (any_bad || S > SMAX) {
            for (int i = tid; i < HZ * 12; i += NT) io.controls[(size_t)inst * HZ * 12 + i] = 0.0;
            if (io.states)
                for (int i = tid; i < HZ * 13; i += NT) io.states[(size_t)inst * HZ * 13 + i] = 0.0;
            if (io.tau)
                for (int i = tid; i < 10; i += NT) io.tau[(size_t)inst * 10 + i] = 0.0;
            if (io.fric_active)
                for (int i = tid; i < HZ; i += NT) io.fric_active[(size_t)inst * HZ + i] = 0;
            if (io.ws_mask)
                for (int i = tid; i < 2 * HZ; i += NT) io.ws_mask[(size_t)inst * 2 * HZ + i] = -1;
            if (tid == 0) {
                io.status[inst] = 3;
                io.iters[inst] = 0;
                if (io.resid) io.resid[2 * inst] = 0.0, io.resid[2 * inst + 1] = 0.0;
            }
            gsync<NT>();
            continue;
        }

        // ---- 2. prefix sums P_k = sum_{l=1..k} Rinv_l, input maps W_j, constraint rows ----
        for (int k = tid; k < HZ; k += NT) {
            double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int l = 1; l <= k; ++l)
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] += rinv[9 * l + i];
#pragma unroll
            for (int i = 0; i < 9; ++i) psum[9 * k + i] = acc[i];
        }
        for (int j = tid; j < S; j += NT) {
            const int s = blk_stage[j], l = blk_foot[j];
            const double* fr = footv + 6 * footsel[s] + 3 * l;
            const double r0 = fr[0] - xref[12 * s + 3], r1 = fr[1] - xref[12 * s + 4], r2 = fr[2] - xref[12 * s + 5];
            const double* ii = iwinv + 9 * s;
            // B_omega = dt * Iw^{-1} [skew(r) | I]   (MPC.py:174-179, 184)
            double B[18];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double i0 = ii[3 * a], i1 = ii[3 * a + 1], i2 = ii[3 * a + 2];
                B[6 * a + 0] = dt * (i1 * r2 - i2 * r1);
                B[6 * a + 1] = dt * (i2 * r0 - i0 * r2);
                B[6 * a + 2] = dt * (i0 * r1 - i1 * r0);
                B[6 * a + 3] = dt * i0;
                B[6 * a + 4] = dt * i1;
                B[6 * a + 5] = dt * i2;
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int c = 0; c < LB; ++c) Wm[(j * 3 + a) * LB + c] = B[6 * a + p.comps[c]];
                double wp = 0.0, vp = 0.0;
                for (int c = 0; c < p.npinned; ++c) {
                    wp += B[6 * a + p.pinned[c]] * p.lo6[p.pinned[c]];
                    if (p.pinned[c] == a) vp += dt / p.mass * p.lo6[p.pinned[c]];
                }
                Wp[3 * j + a] = wp;
                Vp[3 * j + a] = vp;
            }
        }
        // per-block inequality rows in block coordinates (MPC.py:220-271), same for every block
        for (int k = tid; k < mb; k += NT) {
            const int kind = p.row_kind[k], arg = p.row_arg[k];
            double f6[6] = {0, 0, 0, 0, 0, 0};
            double rhs = 0.0;
            if (kind == ROW_LO) {
                f6[p.comps[arg]] = -1.0;
                rhs = -p.lo6[p.comps[arg]];
            } else if (kind == ROW_HI) {
                f6[p.comps[arg]] = 1.0;
                rhs = p.hi6[p.comps[arg]];
            } else if (kind == ROW_FRIC) {
                f6[arg & 1] = (arg < 2) ? 1.0 : -1.0;
                f6[2] = -p.mu;
            } else {
                const double len = (arg == 0) ? p.lh_eff : p.lt_eff;
                const double sg = (arg == 0) ? 1.0 : -1.0;
                // [0,0,1] @ R.T = third column of R;  [0,1,0] @ R.T = second column
                f6[0] = -len * rotn[2], f6[1] = -len * rotn[5], f6[2] = -len * rotn[8];
                f6[3] = sg * rotn[1], f6[4] = sg * rotn[4], f6[5] = sg * rotn[7];
            }
            for (int c = 0; c < p.npinned; ++c) rhs -= f6[p.pinned[c]] * p.lo6[p.pinned[c]];
#pragma unroll
            for (int c = 0; c < LB; ++c) Cb[k * LB + c] = f6[p.comps[c]];
            rb[k] = rhs;
        }
        gsync<NT>();

        // ---- 3. free-response error and the stage-pair kernels ----------------------------
        // e_i = X_i(u = 0, pinned components at their bound) - x_ref_i
        for (int i = tid; i < HZ; i += NT) {
            const double* P = psum + 9 * i;
            const double* R0 = rinv;
            double e[12];
            const double c1 = (double)(i + 1), c2 = 0.5 * (double)i * (double)(i + 1);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double m0 = R0[3 * a] + P[3 * a], m1 = R0[3 * a + 1] + P[3 * a + 1], m2 = R0[3 * a + 2] + P[3 * a + 2];
                e[a] = x_fb[a] + dt * (m0 * x_fb[6] + m1 * x_fb[7] + m2 * x_fb[8]);
                e[3 + a] = x_fb[3 + a] + dt * c1 * x_fb[9 + a];
                e[6 + a] = x_fb[6 + a];
                e[9 + a] = x_fb[9 + a];
            }
            e[5] -= dt * dt * p.g * c2;
            e[11] -= c1 * dt * p.g;
            if (p.npinned > 0) {
                for (int j = 0; j < S; ++j) {
                    const int s = blk_stage[j];
                    if (s > i) break;
                    const double* Ps = psum + 9 * s;
                    const double w0 = Wp[3 * j], w1 = Wp[3 * j + 1], w2 = Wp[3 * j + 2];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        e[a] += dt * ((P[3 * a] - Ps[3 * a]) * w0 + (P[3 * a + 1] - Ps[3 * a + 1]) * w1 +
                                      (P[3 * a + 2] - Ps[3 * a + 2]) * w2);
                        e[3 + a] += dt * (double)(i - s) * Vp[3 * j + a];
                        e[9 + a] += Vp[3 * j + a];
                    }
                    e[6] += w0, e[7] += w1, e[8] += w2;
                }
            }
#pragma unroll
            for (int a = 0; a < 12; ++a) err[12 * i + a] = e[a] - xref[12 * i + a];
        }
        // T(sr,sc)[y][z] = sum_{i>sr} dt^2 (P_i-P_sr)[x][y] Qth[x] (P_i-P_sc)[x][z] + (HZ-sr) Qw[y] d_yz,
        // cpp(sr,sc) = sum_{i>=sr} (i-sr)(i-sc)         (orientation / position coupling of two stages)
        for (int pi = tid; pi < L::NPAIR; pi += NT) {
            int sr = (int)((sqrtf(8.0f * (float)pi + 1.0f) - 1.0f) * 0.5f);
            if (sr * (sr + 1) / 2 > pi) --sr;
            if ((sr + 1) * (sr + 2) / 2 <= pi) ++sr;
            const int sc = pi - sr * (sr + 1) / 2;
            const double* Pr = psum + 9 * sr;
            const double* Pc = psum + 9 * sc;
            double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            double cpp = 0.0;
            for (int i = sr + 1; i < HZ; ++i) {
                cpp += (double)(i - sr) * (double)(i - sc);
                const double* Pi = psum + 9 * i;
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    const double q = p.Q[x] * dt * dt;
#pragma unroll
                    for (int y = 0; y < 3; ++y) {
                        const double dr = q * (Pi[3 * x + y] - Pr[3 * x + y]);
#pragma unroll
                        for (int z = 0; z < 3; ++z) T[3 * y + z] += dr * (Pi[3 * x + z] - Pc[3 * x + z]);
                    }
                }
            }
            const double cnt = (double)(HZ - sr);
            T[0] += cnt * p.Q[6], T[4] += cnt * p.Q[7], T[8] += cnt * p.Q[8];
            double* o = pairs + pi * 10;
#pragma unroll
            for (int x = 0; x < 9; ++x) o[x] = T[x];
            o[9] = cpp;
        }
        gsync<NT>();

        // ---- 4. condensed Hessian tiles and gradient ---------------------------------------
        {
            const double vm = dt / p.mass;
            const int ntile = RIC ? 0 : S * (S + 1) / 2;  // the Riccati backend never forms Hc
            for (int t = tid; t < ntile; t += NT) {
                int jr = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                if (jr * (jr + 1) / 2 > t) --jr;
                if ((jr + 1) * (jr + 2) / 2 <= t) ++jr;
                const int jc = t - jr * (jr + 1) / 2;
                const int sr = blk_stage[jr], sc = blk_stage[jc];  // sc <= sr
                const double* pr = pairs + tidx(sr, sc) * 10;
                const double* Wr = Wm + jr * 3 * LB;
                const double* Wc = Wm + jc * 3 * LB;
                double tmp[3 * LB];
#pragma unroll
                for (int y = 0; y < 3; ++y)
#pragma unroll
                    for (int b = 0; b < LB; ++b)
                        tmp[y * LB + b] = pr[3 * y] * Wc[b] + pr[3 * y + 1] * Wc[LB + b] + pr[3 * y + 2] * Wc[2 * LB + b];
                double blk[E];
#pragma unroll
                for (int a = 0; a < LB; ++a)
#pragma unroll
                    for (int b = 0; b < LB; ++b)
                        blk[a * LB + b] = Wr[a] * tmp[b] + Wr[LB + a] * tmp[LB + b] + Wr[2 * LB + a] * tmp[2 * LB + b];
                const double cpp = pr[9], cnt = (double)(HZ - sr);
                // force components: V = (dt/m) e_a -> diagonal terms only
#pragma unroll
                for (int a = 0; a < LB; ++a) {
                    const int ca = p.comps[a];
                    if (ca < 3) blk[a * LB + a] += vm * vm * (dt * dt * cpp * p.Q[3 + ca] + cnt * p.Q[9 + ca]);
                }
                if (jr == jc) {
                    const int l = blk_foot[jr];
#pragma unroll
                    for (int a = 0; a < LB; ++a) {
                        const int ca = p.comps[a];
                        blk[a * LB + a] += p.R[(ca < 3) ? (3 * l + ca) : (6 + 3 * l + ca - 3)];
                    }
                    // exact symmetry of the stored diagonal tile
#pragma unroll
                    for (int a = 0; a < LB; ++a)
#pragma unroll
                        for (int b = 0; b < a; ++b) blk[b * LB + a] = blk[a * LB + b];
                }
                tile_store<LB>(Mb + t * TS, blk);
            }
            // gradient: one lane per block, taken from the top so they overlap the tile loop
            for (int jj = tid; jj < S; jj += NT) {
                const int j = S - 1 - jj, s = blk_stage[j];
                const double* Wj = Wm + j * 3 * LB;
                double qw[3] = {0, 0, 0}, qv[3] = {0, 0, 0}, qp[3] = {0, 0, 0}, qt[3] = {0, 0, 0};
                for (int i = s; i < HZ; ++i) {
                    const double* e = err + 12 * i;
                    const double* Pi = psum + 9 * i;
                    const double* Ps = psum + 9 * s;
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        qw[x] += p.Q[6 + x] * e[6 + x];
                        qv[x] += p.Q[9 + x] * e[9 + x];
                        qp[x] += dt * (double)(i - s) * p.Q[3 + x] * e[3 + x];
                    }
                    const double t0 = p.Q[0] * e[0], t1 = p.Q[1] * e[1], t2 = p.Q[2] * e[2];
#pragma unroll
                    for (int y = 0; y < 3; ++y)
                        qt[y] += dt * ((Pi[y] - Ps[y]) * t0 + (Pi[3 + y] - Ps[3 + y]) * t1 + (Pi[6 + y] - Ps[6 + y]) * t2);
                }
#pragma unroll
                for (int a = 0; a < LB; ++a) {
                    double acc = 0.0;
#pragma unroll
                    for (int x = 0; x < 3; ++x) acc += Wj[x * LB + a] * (qt[x] + qw[x]);
                    const int ca = p.comps[a];
                    if (ca < 3) acc += vm * (qp[ca] + qv[ca]);
                    gv[j * LB + a] = acc;
                }
            }
        }
        gsync<NT>();
        if constexpr (RIC) {  // problem data the stage-wise backend keeps for the whole solve
            for (int i = tid; i < S * 3 * LB; i += NT) ric.W0()[i] = Wm[i];
            for (int i = tid; i < HZ * 9; i += NT) ric.rinv()[i] = rinv[i];
            for (int i = tid; i < HZ * 12; i += NT) ric.err()[i] = err[i];
        }
        // the assembly arrays are needed again only by the output stage: park them in the scratch
        for (int i = tid; i < L::a_end - L::o_work; i += NT) gsave[i] = sm[L::o_work + i];
        // H -> global scratch (16-byte coalesced stores); TMA brings it back once per iteration
        if constexpr (!RIC) {
            const int nd2 = (S * (S + 1) / 2) * TS / 2;
            const double2* src = reinterpret_cast<const double2*>(Mb);
            double2* dst = reinterpret_cast<double2*>(hglob);
            for (int i = tid; i < nd2; i += NT) dst[i] = src[i];
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        gsync<NT>();  // the work region changes hands: assembly arrays -> solver vectors / row arrays
        const uint32_t hbytes = (uint32_t)((S * (S + 1) / 2) * TS * 8);
        if (!RIC && io.dbg_H != nullptr && w == group && group == 0) {
            const int nmax = 12 * HZ;
            for (int e = tid; e < n * n; e += NT) {
                int i = e / n, j = e - i * n;
                if (j > i) {
                    const int t = i;
                    i = j, j = t;
                }
                io.dbg_H[(e / n) * nmax + (e - (e / n) * n)] = Mb[tidx(i / LB, j / LB) * TS + (i % LB) * LB + (j % LB)];
            }
            for (int i = tid; i < n; i += NT) io.dbg_g[i] = gv[i];
            if (tid == 0) io.dbg_n[0] = n;
        }

        // ---- 5. interior point (Mehrotra predictor-corrector + one Gondzio centrality
        //         corrector) to a loose tolerance, then the active-set polish --------------------
        const float inv_mb = 1.0f / (float)mb;
        // row r = j * mb + k: block j, row kind k (exact for these sizes: (r + .5) / mb is never near an integer)
#define BMPC_FOR_ROWS(r, j, k)                                  \
    _Pragma("unroll 1") for (int r = tid; r < m; r += NT)       \
        for (int j = (int)(((float)r + 0.5f) * inv_mb), k = r - j * mb, once_ = 1; once_; once_ = 0)
        auto cdot = [&](int j, int k, const double* v) {  // (C v)_r
            const double* cb = Cb + k * LB;
            const double* vv = v + j * LB;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < LB; ++c) acc += cb[c] * vv[c];
            return acc;
        };
        auto col_gather = [&](int i, const double* wr) {  // (C' w)_i for variable i
            const int j = i / LB, c = i - j * LB;
            const double* ww = wr + j * mb;
            double acc = 0.0;
#pragma unroll 1
            for (int k = 0; k < mb; ++k) acc += Cb[k * LB + c] * ww[k];
            return acc;
        };
        // H is in Mb (h_valid) or on its way (h_pending)
        bool h_valid = true, h_pending = false;
        auto h_issue = [&]() {  // all threads: make Mb reusable, then one lane issues the bulk copy
            if constexpr (RIC) {  // nothing to reload, but callers rely on the barrier this function implies
                gsync<NT>();
                return;
            }
            if constexpr (MG) {  // matrix in global memory: a plain cooperative copy H -> working copy
                gsync<NT>();
                const int nd2 = (S * (S + 1) / 2) * TS / 2;
                const double2* src = reinterpret_cast<const double2*>(hglob);
                double2* dst = reinterpret_cast<double2*>(Mb);
                for (int i = tid; i < nd2; i += NT) dst[i] = src[i];
                gsync<NT>();
                h_pending = false;
                h_valid = true;
                return;
            }
            asm volatile("fence.proxy.async;" ::: "memory");
            gsync<NT>();
            if (tid == 0) {
                mbar_expect_tx(&bars[2], hbytes);
                tma_load_1d(Mb, hglob, hbytes, &bars[2]);
            }
            h_pending = true;
            h_valid = false;
        };
        auto h_need = [&]() {  // all threads: H usable in Mb after this
            if constexpr (RIC) return;
            if (!h_valid && !h_pending) h_issue();
            if (h_pending) {
                mbar_wait(&bars[2], hparity);
                hparity ^= 1u;
                h_pending = false;
                h_valid = true;
            }
        };

        for (int i = tid; i < n; i += NT) uv[i] = ub[i % LB];
        double part = 0.0;
        BMPC_FOR_ROWS(r, j, k) {
            double sl = rb[k] - cdot(0, k, ub);
            if (!(sl > 1e-3)) sl = 1.0;  // infeasible start for this row: handled through rp
            r_s[r] = sl;
            part += sl;
        }
        int status = 1, it = 0;
        double mu = 0.0, rdmax = 0.0;
        bool polished = false, rd_fresh = false;
        if (n > 0) {
            double gpart = 0.0;
            for (int i = tid; i < n; i += NT) gpart = fmax(gpart, fabs(gv[i]));
            const double gs = 1.0 + gmax<NT>(gpart, red);
            const double mu0 = p.mu0_scale * gsum<NT>(part, red) / (double)m;
            gsync<NT>();

            double mu_target = p.mu_tol * gs;
            // attempt -1 (warm start, closed loop only): skip the interior point and polish from the
            // previous tick's active set shifted by one stage; if that does not certify, the cold
            // path below runs exactly as if no warm start had been given
            const bool warm = io.ws_mask != nullptr && io.warm != 0;
            for (int attempt = warm ? -1 : 0; attempt < 3 && !polished; mu_target *= (attempt >= 0 ? 1e-2 : 1.0), ++attempt) {
                if (attempt == 0) {
                    BMPC_FOR_ROWS(r, j, k) r_l[r] = mu0 / r_s[r];
                    gsync<NT>();
                }
                if constexpr (RIC)
                    if (attempt >= 0) ric_set_maps<LB, NT>(p, ric, S, nullptr);
                // ======================= interior-point iterations =========================
                while (attempt >= 0) {
                    if (it >= p.max_iter) {
                        status = 1;
                        break;
                    }
                    ++it;
                    lock_sync();
                    part = 0.0;
                    BMPC_FOR_ROWS(r, j, k) {
                        const double s = r_s[r], l = r_l[r];
                        const double d = l / s, rp = cdot(j, k, uv) + s - rb[k];
                        r_d[r] = d;
                        r_p[r] = rp;
                        r_w[r] = d * rp - l;  // predictor right-hand side row term (used below)
                        part += s * l;
                    }
                    h_need();  // H back in Mb (its reload was issued right after the last solve of the previous iteration)
                    gsync<NT>();
                    if (!rd_fresh) {
                        // stationarity residual rd = Hc u + g + C' lam, evaluated once; a Newton step of
                        // length alpha scales it by (1 - alpha) exactly, so afterwards it is carried by
                        // that recurrence (the polish re-derives everything exactly anyway)
                        if constexpr (RIC) ric_grad<LB, NT>(p, ric, uv, rdv);
                        else tile_symv<LB, NT>(Mb, S, uv, gv, rdv);
                        gsync<NT>();
                        double rdp = 0.0;
#pragma unroll 1
                        for (int i = tid; i < n; i += NT) {
                            const double v = rdv[i] + col_gather(i, r_l);
                            rdv[i] = v;
                            rdp = fmax(rdp, fabs(v));
                        }
                        rdmax = gmax<NT>(rdp, red);
                        rd_fresh = true;
                    }
                    mu = gsum<NT>(part, red) / (double)m;
                    if (mu <= mu_target && rdmax <= p.rd_tol * mu_target) {
                        status = 0;
                        break;
                    }
                    // M = Hc + blockdiag(Cb' diag(d_j) Cb); predictor rhs = -rd - C'(d rp - lam)
#pragma unroll 1
                    for (int e = tid; e < (RIC ? 0 : S * NAB); e += NT) {
                        const int j = e / NAB, ab = e - j * NAB;
                        int a = 0, b = ab;  // ab = a(a+1)/2 + b
                        while (b > a) ++a, b -= a;
                        const double* dd = r_d + j * mb;
                        double acc = 0.0;
#pragma unroll 1
                        for (int k = 0; k < mb; ++k) acc += Cb[k * LB + a] * Cb[k * LB + b] * dd[k];
                        double* dg = Mb + tidx(j, j) * TS;
                        dg[a * LB + b] += acc;
                        if (a != b) dg[b * LB + a] += acc;
                    }
                    gsync<NT>();
#pragma unroll 1
                    for (int i = tid; i < n; i += NT) xv[i] = -rdv[i] - col_gather(i, r_w);
                    h_valid = false;
                    gsync<NT>();
                    bool fact_ok;
                    if constexpr (RIC) {
                        RtSpec rt;
                        rt.polish = 0, rt.Cb = Cb, rt.dd = r_d, rt.mb = mb, rt.Nn = nullptr, rt.bdim = nullptr;
                        fact_ok = ric_factor<LB, NT>(p, ric, rt, S);
                    } else {
                        fact_ok = tile_factor<LB, NT>(Mb, S);
                    }
                    if (!fact_ok) {
                        status = 2;
                        break;
                    }
                    if constexpr (RIC) ric_solve<LB, NT>(ric, xv);
                    else tile_solve<LB, NT, NPT>(Mb, S, xv);  // xv = du_aff
                    // affine step: ratio test in FP32 (only a step LENGTH, cut by 0.995 afterwards), and
                    // mu_aff = mu (1 - a) + a^2 sum(dsa dla) / m   because  s dla + lam dsa = -s lam
                    float ratio = 0.f;
                    part = 0.0;
                    BMPC_FOR_ROWS(r, j, k) {
                        const double dsa = -r_p[r] - cdot(j, k, xv);
                        const double dla = -r_l[r] - r_d[r] * dsa;
                        ratio = fmaxf(ratio, fmaxf(step_ratio(dsa, r_s[r]), step_ratio(dla, r_l[r])));
                        r_c[r] = dsa * dla;
                        part += dsa * dla;
                    }
#pragma unroll 1
                    for (int i = tid; i < n; i += NT) duv[i] = xv[i];
                    ratio = gmaxf<NT>(ratio, red);
                    const double a_aff = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
                    const double mu_aff = mu * (1.0 - a_aff) + a_aff * a_aff * gsum<NT>(part, red) / (double)m;
                    double sigma = mu_aff / mu;
                    sigma = sigma * sigma * sigma;
                    const double tgt = sigma * mu;
                    // corrector: du = du_aff + inv(M) C' wc,  wc = (dsa dla - sigma mu) / s
                    BMPC_FOR_ROWS(r, j, k) r_c[r] = (r_c[r] - tgt) / r_s[r];
                    gsync<NT>();
#pragma unroll 1
                    for (int i = tid; i < n; i += NT) xv[i] = col_gather(i, r_c);
                    gsync<NT>();
                    if constexpr (RIC) ric_solve<LB, NT>(ric, xv);
                    else tile_solve<LB, NT, NPT>(Mb, S, xv);
#pragma unroll 1
                    for (int i = tid; i < n; i += NT) duv[i] += xv[i];
                    gsync<NT>();
                    // rows of the combined direction: ds, dlam kept for the passes below
                    ratio = 0.f;
                    BMPC_FOR_ROWS(r, j, k) {
                        const double ds = -r_p[r] - cdot(j, k, duv);
                        const double dl = -r_l[r] - r_c[r] - r_d[r] * ds;
                        ratio = fmaxf(ratio, fmaxf(step_ratio(ds, r_s[r]), step_ratio(dl, r_l[r])));
                        r_p[r] = ds;  // rp is not needed any more this iteration
                        r_c[r] = dl;
                    }
                    ratio = gmaxf<NT>(ratio, red);
                    double a2 = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
                    if (!isfinite(ratio)) {
                        status = 2;
                        break;
                    }
                    // Gondzio centrality corrector: pull outlier products of the trial point into
                    // [0.1, 10] x target; one more pair of triangular solves with the same factor.
                    // Direction change: du += x, ds -= C x, dlam += d (C x) - w,  w = -vt / s
                    if (p.gondzio && a2 < p.gondzio_below) {
                        const double at = fmin(1.0, 1.5 * a2 + 0.1);
                        BMPC_FOR_ROWS(r, j, k) {
                            const double v = (r_s[r] + at * r_p[r]) * (r_l[r] + at * r_c[r]);
                            double vt = fmin(fmax(v, 0.1 * tgt), 10.0 * tgt) - v;
                            vt = fmax(vt, -10.0 * tgt);
                            r_w[r] = -vt / r_s[r];
                        }
                        gsync<NT>();
#pragma unroll 1
                        for (int i = tid; i < n; i += NT) xv[i] = col_gather(i, r_w);
                        gsync<NT>();
                        if constexpr (RIC) ric_solve<LB, NT>(ric, xv);
                        else tile_solve<LB, NT, NPT>(Mb, S, xv);
                        h_issue();  // the factor is dead: bring H back while the step length and the update are computed
                        ratio = 0.f;
                        BMPC_FOR_ROWS(r, j, k) {
                            const double cx = cdot(j, k, xv);
                            const double ds = r_p[r] - cx;
                            const double dl = r_c[r] + r_d[r] * cx - r_w[r];
                            ratio = fmaxf(ratio, fmaxf(step_ratio(ds, r_s[r]), step_ratio(dl, r_l[r])));
                        }
                        ratio = gmaxf<NT>(ratio, red);
                        const double a3 = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
                        if (isfinite(ratio) && a3 > a2) {
                            a2 = a3;
                            BMPC_FOR_ROWS(r, j, k) {
                                const double cx = cdot(j, k, xv);
                                r_p[r] -= cx;
                                r_c[r] += r_d[r] * cx - r_w[r];
                            }
#pragma unroll 1
                            for (int i = tid; i < n; i += NT) duv[i] += xv[i];
                        }
                    }
                    const double alpha = (it > 14 ? 0.9 : p.step_frac) * a2;  // late iterations: stay well inside (anti-stall)
                    BMPC_FOR_ROWS(r, j, k) {
                        r_s[r] += alpha * r_p[r];
                        r_l[r] += alpha * r_c[r];
                    }
#pragma unroll 1
                    for (int i = tid; i < n; i += NT) uv[i] += alpha * duv[i], rdv[i] *= (1.0 - alpha);
                    rdmax *= (1.0 - alpha);
                    if (!h_pending && !h_valid) h_issue();  // (the Gondzio branch issues it earlier when it runs)
                    gsync<NT>();
                }

                // ============================ active-set polish ============================
                if (attempt < 0) {
                    // guess: the certified active set of the previous tick, one stage later (same absolute
                    // time); the new last stage copies the old last stage.  No multiplier estimate: lam = 0
                    const int32_t* wm = io.ws_mask + (size_t)inst * 2 * HZ;
                    for (int j = tid; j < S; j += NT) {
                        const int s = blk_stage[j], l = blk_foot[j], s1 = min(s + 1, HZ - 1);
                        int mk = wm[2 * s1 + l];
                        if (mk < 0) mk = wm[2 * s1 + (1 - l)];
                        if (mk < 0) mk = wm[2 * s + l];
                        amask[j] = mk < 0 ? 0 : mk;
                    }
                    BMPC_FOR_ROWS(r, j, k) r_l[r] = 0.0;
                    gsync<NT>();
                } else {
                // guess: row active when its barrier weight lam/s dominates the curvature along it
                for (int j = tid; j < S; j += NT) amask[j] = 0;
                h_need();  // diag(Hc) is read from the tile matrix
                if constexpr (RIC) ric_hdiag<LB, NT>(p, ric, upv);  // ... or evaluated into the (dead) polished-point vector
                gsync<NT>();
                BMPC_FOR_ROWS(r, j, k) {
                    const double* cb = Cb + k * LB;
                    const double* hh = RIC ? upv + j * LB : Mb + tidx(j, j) * TS;
                    constexpr int hstep = RIC ? 1 : LB + 1;
                    double th = 0.0, aa = 0.0;
#pragma unroll
                    for (int c = 0; c < LB; ++c) th += cb[c] * cb[c] * hh[c * hstep], aa += cb[c] * cb[c];
                    if (r_l[r] * fmax(aa * aa, 1e-300) > th * r_s[r]) atomicOr(&amask[j], 1 << k);
                }
                gsync<NT>();
                }
                bool ok = false;
                // the last attempt may take six times the rounds: a nearly degenerate instance whose first active-set guesses are
                // far off (one in ~65,000 at h = 30) walks to the optimal set one row per block and round (see the release rule below)
                const int max_rounds = attempt < 0 ? p.warm_rounds : (attempt == 2 ? 6 * p.polish_rounds : p.polish_rounds);
                for (int round = 0; round < max_rounds; ++round) {
                    lock_sync();
                    // per block: affine set of the active rows  u_b = p_b + N_b w_b
                    int bad_blk = 0;
                    for (int j = tid; j < S; j += NT) {
                        int dim = 0;
                        if (!block_nullspace<LB>(Cb, rb, mb, (unsigned)amask[j], ppv + j * LB, Nn + j * E, &dim)) bad_blk = 1;
                        bdim[j] = dim;
                    }
                    if (gany<NT>(bad_blk)) break;
                    gsync<NT>();
                    h_need();
                    if constexpr (RIC) ric_grad<LB, NT>(p, ric, ppv, tvp);
                    else tile_symv<LB, NT>(Mb, S, ppv, gv, tvp);  // Hc p + g
                    gsync<NT>();
                    // reduced system, padded to the tile grid: tile <- N_jr' H N_jc (+ I on the padding)
                    for (int i = tid; i < n; i += NT) {
                        const int j = i / LB, a = i - j * LB;
                        double acc = 0.0;
#pragma unroll
                        for (int c = 0; c < LB; ++c) acc += Nn[j * E + c * LB + a] * tvp[j * LB + c];
                        xv[i] = (a < bdim[j]) ? -acc : 0.0;
                    }
                    if constexpr (!RIC) {
                        const int ntile = S * (S + 1) / 2;
                        for (int t = tid; t < ntile; t += NT) {
                            int jr = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                            if (jr * (jr + 1) / 2 > t) --jr;
                            if ((jr + 1) * (jr + 2) / 2 <= t) ++jr;
                            const int jc = t - jr * (jr + 1) / 2;
                            double hh[E], tmp[E];
                            tile_load<LB>(Mb + t * TS, hh);
                            const double* nc = Nn + jc * E;
                            const double* nr = Nn + jr * E;
#pragma unroll
                            for (int a = 0; a < LB; ++a)
#pragma unroll
                                for (int b = 0; b < LB; ++b) {
                                    double acc = 0.0;
#pragma unroll
                                    for (int c = 0; c < LB; ++c) acc += hh[a * LB + c] * nc[c * LB + b];
                                    tmp[a * LB + b] = acc;
                                }
#pragma unroll
                            for (int a = 0; a < LB; ++a)
#pragma unroll
                                for (int b = 0; b < LB; ++b) {
                                    double acc = 0.0;
#pragma unroll
                                    for (int c = 0; c < LB; ++c) acc += nr[c * LB + a] * tmp[c * LB + b];
                                    hh[a * LB + b] = acc;
                                }
                            if (jr == jc) {
                                const int dim = bdim[jr];
#pragma unroll
                                for (int a = 0; a < LB; ++a)
                                    if (a >= dim) hh[a * LB + a] = 1.0;
                            }
                            tile_store<LB>(Mb + t * TS, hh);
                        }
                    }
                    h_valid = false;
                    gsync<NT>();
                    if constexpr (RIC) {
                        // the same LQR with inputs w_b: B_b N_b in place of B_b, N_b' R N_b (+ I on the padding) as weights
                        ric_set_maps<LB, NT>(p, ric, S, Nn);
                        RtSpec rt;
                        rt.polish = 1, rt.Cb = Cb, rt.dd = nullptr, rt.mb = mb, rt.Nn = Nn, rt.bdim = bdim;
                        if (!ric_factor<LB, NT>(p, ric, rt, S)) break;
                        ric_solve<LB, NT>(ric, xv);
                    } else {
                        if (!tile_factor<LB, NT>(Mb, S)) break;
                        tile_solve<LB, NT, NPT>(Mb, S, xv);
                    }
                    for (int i = tid; i < n; i += NT) {
                        const int j = i / LB, c = i - j * LB;
                        double acc = ppv[i];
#pragma unroll
                        for (int a = 0; a < LB; ++a) acc += Nn[j * E + c * LB + a] * xv[j * LB + a];
                        upv[i] = acc;
                    }
                    h_issue();  // (syncs) H needed again for the multiplier check
                    // primal check: violated inactive rows join the active set
                    int changed = 0;
                    BMPC_FOR_ROWS(r, j, k) {
                        const double bk = rb[k];
                        const double viol = cdot(j, k, upv) - bk;
                        if (viol > 1e-9 * (1.0 + fabs(bk)) && !((amask[j] >> k) & 1)) {
                            atomicOr(&amask[j], 1 << k);
                            changed = 1;
                        }
                    }
                    if (gany<NT>(changed)) {
                        gsync<NT>();
                        continue;
                    }
                    // dual check: minus the gradient must be a non-negative combination of the active rows
                    h_need();
                    if constexpr (RIC) ric_grad<LB, NT>(p, ric, upv, tvp);
                    else tile_symv<LB, NT>(Mb, S, upv, gv, tvp);
                    gsync<NT>();
                    int fail = 0;
                    for (int j = tid; j < S; j += NT) {
                        double rneg[LB];
#pragma unroll
                        for (int c = 0; c < LB; ++c) rneg[c] = -tvp[j * LB + c];
                        unsigned drop = 0u;
                        if (block_dual_fast<LB>(Cb, mb, (unsigned)amask[j], r_l + j * mb, rneg, gs)) continue;
                        if (!block_dual_check<LB>(Cb, mb, (unsigned)amask[j], rneg, gs, &drop)) {
                            // releasing several rows of many blocks at once can cycle (release, re-add as violated, release ...):
                            // after the first rounds only one row per block is released at a time
                            if (round > 3 && drop != 0u) drop &= (~drop + 1u);
                            if (drop == 0u) fail = 1;
                            else amask[j] &= ~(int)drop, changed = 1;
                        }
                    }
                    const int any_fail = gany<NT>(fail);
                    const int any_changed = gany<NT>(changed);
                    gsync<NT>();
                    if (any_fail) break;
                    if (!any_changed) {
                        ok = true;
                        break;
                    }
                }
                if (ok) {
                    polished = true;
                    status = 0;
                    for (int i = tid; i < n; i += NT) uv[i] = upv[i];
                    gsync<NT>();
                } else if (status == 2 || it >= p.max_iter) {
                    break;  // cannot iterate further: return the interior-point iterate
                } else {
                    status = 1;
                }
            }
            if (h_pending) h_need();  // never leave a bulk copy in flight into Mb
        } else {
            status = 0;
        }
#undef BMPC_FOR_ROWS
        gsync<NT>();

        // ---- 6. outputs ----------------------------------------------------------------------
        lock_sync();
        // bring the assembly arrays back into the (now free) work region
        for (int i = tid; i < L::a_end - L::o_work; i += NT) sm[L::o_work + i] = gsave[i];
        gsync<NT>();
        // controls (h,12): swing feet 0, pinned components at their bound (MPC.py:300-302)
        double umax_part = 0.0;
        for (int e = tid; e < HZ * 12; e += NT) {
            const int s = e / 12, c12 = e - 12 * s;
            const int l = (c12 % 6) / 3, comp = (c12 < 6) ? (c12 % 3) : (3 + c12 % 3);
            const int b = blockOf[2 * s + l];
            double val = 0.0;
            if (b >= 0) {
                val = p.lo6[comp];
#pragma unroll
                for (int c = 0; c < LB; ++c)
                    if (p.comps[c] == comp) val = uv[b * LB + c];
            }
            io.controls[(size_t)inst * HZ * 12 + e] = val;
            umax_part = fmax(umax_part, fabs(val));
            if (s == 0) gv[c12] = val;  // first-stage input for the torque map (the gradient is dead now)
        }
        const double uscale = fmax(1.0, gmax<NT>(umax_part, red));
        // predicted states (h,13): X_i = free response + sum_j dX_i/du_j u_j
        if (io.states) {
            for (int i = tid; i < HZ; i += NT) {
                double X[12];
#pragma unroll
                for (int a = 0; a < 12; ++a) X[a] = err[12 * i + a] + xref[12 * i + a];
                const double* Pi = psum + 9 * i;
                for (int j = 0; j < S; ++j) {
                    const int s = blk_stage[j];
                    if (s > i) break;
                    const double* Wj = Wm + j * 3 * LB;
                    double wv[3] = {0, 0, 0}, vv[3] = {0, 0, 0};
#pragma unroll
                    for (int c = 0; c < LB; ++c) {
                        const double uc = uv[j * LB + c];
#pragma unroll
                        for (int x = 0; x < 3; ++x) wv[x] += Wj[x * LB + c] * uc;
                        if (p.comps[c] < 3) vv[p.comps[c]] += dt / p.mass * uc;
                    }
                    const double* Ps = psum + 9 * s;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        X[a] += dt * ((Pi[3 * a] - Ps[3 * a]) * wv[0] + (Pi[3 * a + 1] - Ps[3 * a + 1]) * wv[1] +
                                      (Pi[3 * a + 2] - Ps[3 * a + 2]) * wv[2]);
                        X[3 + a] += dt * (double)(i - s) * vv[a];
                        X[6 + a] += wv[a];
                        X[9 + a] += vv[a];
                    }
                }
                double* out = io.states + ((size_t)inst * HZ + i) * 13;
#pragma unroll
                for (int a = 0; a < 12; ++a) out[a] = X[a];
                out[12] = 1.0;
            }
        }
        // non-trivially active friction rows per stage
        if (io.fric_active) {
            for (int s = tid; s < HZ; s += NT) {
                const double tol = 1e-6 * uscale;
                unsigned mask = 0;
                for (int l = 0; l < 2; ++l) {
                    const int b = blockOf[2 * s + l];
                    if (b < 0) continue;
                    double f[3] = {p.lo6[0], p.lo6[1], p.lo6[2]};
#pragma unroll
                    for (int c = 0; c < LB; ++c)
                        if (p.comps[c] < 3) f[p.comps[c]] = uv[b * LB + c];
                    if (f[2] <= tol) continue;
                    const double res[4] = {f[0] - p.mu * f[2], f[1] - p.mu * f[2], -f[0] - p.mu * f[2], -f[1] - p.mu * f[2]};
                    for (int r = 0; r < 4; ++r)
                        if (res[r] >= -tol) mask |= 1u << (4 * l + r);
                }
                io.fric_active[(size_t)inst * HZ + s] = (uint8_t)mask;
            }
        }
        gsync<NT>();
        // joint torques from the first-stage input (MPC.py:444-470), one lane per leg
        if (io.do_lowlevel && io.tau && tid < 2) {
            const int leg = tid;
            double tl[5];
            lowlevel_leg(p, x_fb, io.t_swing[inst], cur + 38, cur + 18, cur + 28, rotn, leg, (double)cont[leg], gv, tl);
#pragma unroll
            for (int c = 0; c < 5; ++c) io.tau[(size_t)inst * 10 + 5 * leg + c] = tl[c];
        }
        if (io.ws_mask) {  // certified active set per (stage, foot) slot for the next tick's warm start
            for (int e = tid; e < 2 * HZ; e += NT) {
                const int b = blockOf[e];
                io.ws_mask[(size_t)inst * 2 * HZ + e] = (polished && b >= 0) ? amask[b] : -1;
            }
        }
        if (tid == NT - 1) {
            io.status[inst] = status;
            io.iters[inst] = it;
            if (io.resid) io.resid[2 * inst] = polished ? 0.0 : mu, io.resid[2 * inst + 1] = rdmax;
        }
        gsync<NT>();
    }
    lock_drop();
}

}  // namespace bmpc
