// Batched HECTOR-style force-and-moment MPC for sm_100a: device code.
//
// One CTA solves one robot's MPC tick end to end:
//   1. stage inputs into shared memory with TMA bulk copies (cp.async.bulk + mbarrier),
//      prefetching the next instance while the current one is solved;
//   2. build the discretised single-rigid-body model over the horizon and the
//      CONTACT-REDUCED CONDENSED QP  min 1/2 u'Hc u + g'u,  Cb u_b <= rb per block
//      (reference: MPC.py:61-109 references, 148-185 dynamics, 202-286 QP data);
//   3. solve it with a Mehrotra predictor-corrector interior-point method in FP64,
//      the packed Cholesky factor of  Hc + C' diag(lam/s) C  resident in shared memory
//      (replaces cvxopt.solvers.qp, MPC.py:289-297);
//   4. map the first-stage forces/moments to joint torques (MPC.py:306-365, 426-470).
//
// Why these choices (measured): the QP is strictly convex but nearly flat
// in the forces (R = 1e-4 vs Q up to 700, cond(Hc) ~ 1e6), so first-order splitting
// (ADMM/OSQP) needs thousands of iterations to place the forces, while the interior
// point method reaches 1e-7 relative in ~13 iterations.  Every instance has its own
// 50..120-variable matrix and a single right-hand side, so there is no GEMM to put on
// tcgen05; the work is FP64 CUDA-core FMA fed from shared memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bmpc {

constexpr int MAXROWS = 18;  // per-block inequality rows: 2*6 bounds + 4 friction + 2 line-foot
constexpr int IN_DOUBLES = 48;  // x_fb 12 | foot 6 | q 10 | qd 10 | pf_w 6 | pad 4

enum RowKind { ROW_LO = 0, ROW_HI = 1, ROW_FRIC = 2, ROW_LINE = 3 };

struct DevParams {
    int h, extend, LB, mb, npinned, max_iter, gondzio;
    int comps[6];
    int pinned[6];
    int row_kind[MAXROWS];
    int row_arg[MAXROWS];
    double dt, kv, swing_height, mass, lt_eff, lh_eff, g, mu, mu_tol, rd_tol, init_fz_frac;
    double x_cmd[12], Q[13], R[12], kp[9], kd[9], inertia[9], hip[3], lo6[6], hi6[6];
};

struct IoPtrs {
    const double* x_fb;      // [N,12]
    const int32_t* phase_k;  // [N]
    const double* t_swing;   // [N]
    const double* foot;      // [N,6]
    const uint8_t* contact;  // [N,h,2]
    const double* q;         // [N,10]
    const double* qd;        // [N,10]
    const double* pf_w;      // [N,6]
    double* controls;        // [N,h,12]
    double* states;          // [N,h,13] or null
    double* tau;             // [N,10] or null
    int32_t* status;         // [N]
    int32_t* iters;          // [N]
    uint8_t* fric_active;    // [N,h] or null
    double* resid;           // [N,2] or null
    double* dbg_H;           // debug: dense Hc [nmax*nmax] of work item 0, or null
    double* dbg_g;
    int32_t* dbg_n;
    int use_tma;             // inputs are 16-byte aligned: stage them with cp.async.bulk
    int do_lowlevel;         // q/qd/pf_w/t_swing valid, write tau
};

// ------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int tri(int i, int j) { return (i * (i + 1)) / 2 + j; }  // j <= i

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    __syncthreads();
    return t;
}
template <int NT>
__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) t = fmax(t, red[w]);
    __syncthreads();
    return t;
}

__device__ __forceinline__ void mat3_mul(const double* a, const double* b, double* c) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
__device__ __forceinline__ void mat3_tmul(const double* a, const double* b, double* c) {  // a' * b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
__device__ __forceinline__ bool mat3_inv(const double* a, double* r) {
    double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
    double id = 1.0 / det;
    r[0] = c0 * id;
    r[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    r[3] = c1 * id;
    r[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    r[6] = c2 * id;
    r[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return isfinite(id);
}
// eul2rotm of MPC.py:111-138: Rz(e[2]) Ry(e[1]) Rx(e[0])
__device__ __forceinline__ void eul2rotm(const double* e, double* R) {
    double sr, cr, sp, cp, sy, cy;
    sincos(e[0], &sr, &cr);
    sincos(e[1], &sp, &cp);
    sincos(e[2], &sy, &cy);
    R[0] = cy * cp;
    R[1] = cy * sp * sr - sy * cr;
    R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp;
    R[4] = sy * sp * sr + cy * cr;
    R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;
    R[7] = cp * sr;
    R[8] = cp * cr;
}

// ------------------------------------------------------------------------------------
// shared-memory layout (offsets in doubles), sized on the host with the same formulas
// ------------------------------------------------------------------------------------
template <int HZ, int SMAX, int LB>
struct Layout {
    static constexpr int N = LB * SMAX;              // reduced variables (max)
    static constexpr int HP = N * (N + 1) / 2;       // packed lower triangle of Hc
    static constexpr int MP = (N + 1) * (N + 2) / 2; // packed factor incl. the augmented rhs row
    static constexpr int MR = MAXROWS * SMAX;
    static constexpr int NV = ((N + 1) + 3) & ~3;    // padded vector length
    static constexpr int o_in = 0;                   // 2 x IN_DOUBLES (TMA destinations)
    static constexpr int o_cur = o_in + 2 * IN_DOUBLES;
    static constexpr int o_xref = o_cur + IN_DOUBLES;
    static constexpr int o_rinv = o_xref + HZ * 12;
    static constexpr int o_psum = o_rinv + HZ * 9;
    static constexpr int o_iwinv = o_psum + HZ * 9;
    static constexpr int o_err = o_iwinv + HZ * 9;
    static constexpr int o_footv = o_err + HZ * 12;   // 3 variants x 6
    static constexpr int o_rot = o_footv + 24;        // current-orientation rotation (9) + pad
    static constexpr int o_W = o_rot + 12;            // SMAX x 3 x LB
    static constexpr int o_Wp = o_W + SMAX * 3 * LB;  // SMAX x 3 (pinned components -> omega)
    static constexpr int o_Vp = o_Wp + SMAX * 3;      // SMAX x 3 (pinned components -> velocity)
    static constexpr int o_Cb = o_Vp + SMAX * 3;      // MAXROWS x LB
    static constexpr int o_rb = o_Cb + MAXROWS * LB;  // MAXROWS
    static constexpr int o_ub = o_rb + MAXROWS + 2;   // LB start point (+pad)
    static constexpr int o_H = o_ub + 8;
    static constexpr int o_M = o_H + HP;
    static constexpr int o_g = o_M + MP;
    static constexpr int o_u = o_g + NV;
    static constexpr int o_du = o_u + NV;
    static constexpr int o_x = o_du + NV;    // rhs / solution vector of the triangular solves
    static constexpr int o_t = o_x + NV;     // H*u
    static constexpr int o_inv = o_t + NV;   // 1 / L_kk
    static constexpr int o_wrow = o_inv + NV;  // row vector being gathered by C'
    static constexpr int o_drow = o_wrow + MR; // lam / s
    static constexpr int o_red = o_drow + MR;  // reduction scratch (32) + misc scalars (32)
    static constexpr int o_int = o_red + 64;   // ints from here
    static constexpr int n_int = 2 * SMAX + 2 * HZ + HZ + 2 * HZ + 16;  // blk_stage, blk_foot, blockOf, footsel, contact
    static constexpr int o_bar = o_int + (n_int + 1) / 2 + 2;  // two mbarriers (8-byte aligned)
    static constexpr int total_doubles = o_bar + 2;
    static constexpr size_t bytes = size_t(total_doubles) * 8;
};

// ------------------------------------------------------------------------------------
// packed Cholesky of the (n+1)-row augmented matrix: rows 0..n-1 are M (lower, packed),
// row n holds a right-hand side; after the call row n holds L^{-1} rhs (forward
// substitution for free) and inv[k] = 1/L_kk.  Returns false on a non-positive pivot.
// ------------------------------------------------------------------------------------
template <int NT>
__device__ bool chol_aug(double* M, int n, double* inv) {
    constexpr int TX = 8, TY = NT / 8;
    const int tid = threadIdx.x;
    const int tx = tid & (TX - 1), ty = tid / TX;
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        const double mkk = M[tri(k, k)];
        if (!(mkk > 0.0) || !isfinite(mkk)) {  // uniform: every thread reads the same value
            ok = false;
            break;
        }
        const double ik = rsqrt(mkk);
        for (int i = k + 1 + tid; i <= n; i += NT) M[tri(i, k)] *= ik;
        if (tid == 0) inv[k] = ik;
        __syncthreads();
        for (int i = k + 1 + ty; i <= n; i += TY) {
            const double lik = M[tri(i, k)];
            double* row = M + tri(i, 0);
            const int jend = (i < n) ? i : n - 1;  // augmented row has no diagonal
            for (int j = k + 1 + tx; j <= jend; j += TX) row[j] -= lik * M[tri(j, k)];
        }
        __syncthreads();
    }
    return ok;
}

// x <- L^{-T} x   (x has n entries, one per thread; block-wide, one barrier per step)
template <int NT>
__device__ void solve_backward(const double* M, int n, const double* inv, double* x) {
    const int tid = threadIdx.x;
    double bi = (tid < n) ? x[tid] : 0.0;
    for (int k = n - 1; k >= 0; --k) {
        if (tid == k) {
            bi *= inv[k];
            x[k] = bi;
        }
        __syncthreads();
        if (tid < k) bi -= M[tri(k, tid)] * x[k];
    }
    __syncthreads();
}
// x <- L^{-1} x
template <int NT>
__device__ void solve_forward(const double* M, int n, const double* inv, double* x) {
    const int tid = threadIdx.x;
    double bi = (tid < n) ? x[tid] : 0.0;
    for (int k = 0; k < n; ++k) {
        if (tid == k) {
            bi *= inv[k];
            x[k] = bi;
        }
        __syncthreads();
        if (tid > k && tid < n) bi -= M[tri(tid, k)] * x[k];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------
// leg Jacobian (MPC.py:306-365) and low-level torque map (MPC.py:426-470) for one leg
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void leg_jacobian(const double* q, double side, double* J /*6x5 row-major*/) {
    double s0, c0, s1, c1, s2, c2, s23, c23, s234, c234;
    sincos(q[0], &s0, &c0);
    sincos(q[1], &s1, &c1);
    sincos(q[2], &s2, &c2);
    sincos(q[2] + q[3], &s23, &c23);
    sincos(q[2] + q[3] + q[4], &s234, &c234);
    const double sreach[3] = {0.04 * s234 + 0.22 * s23 + 0.22 * s2, 0.04 * s234 + 0.22 * s23, 0.04 * s234};
    const double creach[3] = {0.04 * c234 + 0.22 * c23 + 0.22 * c2, 0.04 * c234 + 0.22 * c23, 0.04 * c234};
    const double lat = 0.018 * side + 0.0025;
#pragma unroll
    for (int i = 0; i < 30; ++i) J[i] = 0.0;
    const double across = 0.015 * side + c1 * lat - s1 * creach[0];
    const double along = sreach[0] + 0.0135;
    J[0 * 5 + 0] = s0 * along + c0 * across;
    J[1 * 5 + 0] = s0 * across - c0 * along;
    J[5 * 5 + 0] = 1.0;
    const double swing = s1 * lat + c1 * creach[0];
    J[0 * 5 + 1] = -s0 * swing;
    J[1 * 5 + 1] = c0 * swing;
    J[2 * 5 + 1] = s1 * creach[0] - c1 * lat;
    J[3 * 5 + 1] = c0;
    J[4 * 5 + 1] = s0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        J[0 * 5 + 2 + c] = s0 * s1 * sreach[c] - c0 * creach[c];
        J[1 * 5 + 2 + c] = -s0 * creach[c] - c0 * s1 * sreach[c];
        J[2 * 5 + 2 + c] = c1 * sreach[c];
        J[3 * 5 + 2 + c] = -c1 * s0;
        J[4 * 5 + 2 + c] = c0 * c1;
        J[5 * 5 + 2 + c] = s1;
    }
}

// tau_leg[5] for one leg.  R = eul2rotm(x_fb[0:3]); u = [f1,f2,m1,m2] first-stage input.
__device__ __noinline__ void lowlevel_leg(const DevParams& p, const double* x_fb, double t, const double* pf_w,
                                    const double* q, const double* qd, const double* R, int leg, double c_leg,
                                    const double* u, double* tau_leg) {
    const double side = (leg == 0) ? 1.0 : -1.0;
    double J[30];
    leg_jacobian(q + 5 * leg, side, J);
    // foot velocity in "world": R' * Jf * qd   (MPC.py:461)
    double vb[3], vf[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 5; ++c) acc += J[a * 5 + c] * qd[5 * leg + c];
        vb[a] = acc;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) vf[a] = R[a] * vb[0] + R[3 + a] * vb[1] + R[6 + a] * vb[2];
    // swing PD (MPC.py:426-442)
    const double hh = (double)p.h;
    double des[3];
    des[0] = x_fb[3] + x_fb[9] * 1 / 2 * hh / 2 * p.dt + p.kv * (x_fb[3] - p.x_cmd[3]);
    des[1] = x_fb[4] + x_fb[10] * 1 / 2 * hh / 2 * p.dt + p.kv * (x_fb[4] - p.x_cmd[4]) + 0.04 * side;
    const double half = p.dt * hh / 2;
    double tm = fmod(t, half);  // np.remainder: result has the sign of the divisor
    if (tm != 0.0 && ((tm < 0.0) != (half < 0.0))) tm += half;
    des[2] = p.swing_height * sin(3.141592653589793 * tm / half);
    double ep[3], ev[3], fs[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        ep[a] = des[a] - pf_w[3 * leg + a];
        ev[a] = 0.0 - vf[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        fs[a] = (p.kp[3 * a] * ep[0] + p.kp[3 * a + 1] * ep[1] + p.kp[3 * a + 2] * ep[2]) +
                (p.kd[3 * a] * ev[0] + p.kd[3 * a + 1] * ev[1] + p.kd[3 * a + 2] * ev[2]);
    // stance wrench: -[R' f; R' m]   (MPC.py:465)
    double wr[6];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        wr[a] = -(R[a] * u[3 * leg] + R[3 + a] * u[3 * leg + 1] + R[6 + a] * u[3 * leg + 2]);
        wr[3 + a] = -(R[a] * u[6 + 3 * leg] + R[3 + a] * u[6 + 3 * leg + 1] + R[6 + a] * u[6 + 3 * leg + 2]);
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        double st = 0.0, sw = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) st += J[a * 5 + c] * wr[a];
#pragma unroll
        for (int a = 0; a < 3; ++a) sw += J[a * 5 + c] * fs[a];
        tau_leg[c] = st * c_leg + sw * -(c_leg - 1.0);
    }
}

// closed-form foot position in the hip frame (MPC.py:367-404)
__device__ inline void foot_body(const double* q, double side, double* pf) {
    double s0, c0, s1, c1, s2, c2, s3, c3, s4, c4;
    sincos(q[0], &s0, &c0);
    sincos(q[1], &s1, &c1);
    sincos(q[2], &s2, &c2);
    sincos(q[3], &s3, &c3);
    sincos(q[4], &s4, &c4);
    const double fwd_c = c0 * c2 - s0 * s1 * s2, fwd_s = c0 * s2 + c2 * s0 * s1;
    const double lat_c = c2 * s0 + c0 * s1 * s2, lat_s = s0 * s2 - c0 * c2 * s1;
    pf[0] = -(3 * c0) / 200 - (9 * s4 * (c3 * fwd_c - s3 * fwd_s)) / 250 - (11 * c0 * s2) / 50 - (side * s0) / 50 -
            (11 * c3 * fwd_s) / 50 - (11 * s3 * fwd_c) / 50 - (9 * c4 * (c3 * fwd_s + s3 * fwd_c)) / 250 -
            (23 * c1 * side * s0) / 1000 - (11 * c2 * s0 * s1) / 50;
    pf[1] = (c0 * side) / 50 - (9 * s4 * (c3 * lat_c - s3 * lat_s)) / 250 - (3 * s0) / 200 - (11 * s0 * s2) / 50 -
            (11 * c3 * lat_s) / 50 - (11 * s3 * lat_c) / 50 - (9 * c4 * (c3 * lat_s + s3 * lat_c)) / 250 +
            (23 * c0 * c1 * side) / 1000 + (11 * c0 * c2 * s1) / 50;
    pf[2] = (23 * side * s1) / 1000 - (11 * c1 * c2) / 50 - (9 * c4 * (c1 * c2 * c3 - c1 * s2 * s3)) / 250 +
            (9 * s4 * (c1 * c2 * s3 + c1 * c3 * s2)) / 250 - (11 * c1 * c2 * c3) / 50 + (11 * c1 * s2 * s3) / 50 -
            3.0 / 50.0;
}

}  // namespace bmpc
