// Batched HECTOR-style force-and-moment MPC for sm_100a: shared device definitions.
//
// Parameter / pointer structs passed to the kernels, TMA + mbarrier helpers, small 3x3 algebra,
// and the closed-form kinematics of the reference (rotation MPC.py:111-138, leg Jacobian
// MPC.py:306-365, swing PD + torque map MPC.py:426-470, foot position MPC.py:367-404).
// The fused tick kernel is in bmpc_tick.cuh, its active-set polish in bmpc_polish.cuh.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define BMPC_HD __host__ __device__

namespace bmpc {

constexpr int MAXROWS = 18;  // per-block inequality rows: 2*6 bounds + 4 friction + 2 line-foot
constexpr int IN_DOUBLES = 48;  // x_fb 12 | foot 6 | q 10 | qd 10 | pf_w 6 | pad 4

enum RowKind { ROW_LO = 0, ROW_HI = 1, ROW_FRIC = 2, ROW_LINE = 3 };

struct DevParams {
    int h, extend, LB, mb, npinned, max_iter, gondzio, warm_rounds, polish_rounds, lock_mode;
    int lane_prefetch, lane_sync;  // lane-per-robot kernels: bulk L2 prefetch of the next block record; lockstep of the warps of a CTA (bmpc_set_option)
    int comps[6];
    int pinned[6];
    int row_kind[MAXROWS];
    int row_arg[MAXROWS];
    double dt, kv, swing_height, mass, lt_eff, lh_eff, g, mu, mu_tol, rd_tol, init_fz_frac, mu0_scale, step_frac, gondzio_below;
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
    int32_t* ws_mask;        // [N,2h] warm-start store: certified active-row mask per (stage, foot), -1 = none; or null
    int warm;                // use ws_mask as the starting active set (it is always written when non-null)
    int use_tma;             // inputs are 16-byte aligned: stage them with cp.async.bulk
    int do_lowlevel;         // q/qd/pf_w/t_swing valid, write tau
};

// ------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------
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

BMPC_HD __forceinline__ void mat3_mul(const double* a, const double* b, double* c) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
BMPC_HD __forceinline__ void mat3_tmul(const double* a, const double* b, double* c) {  // a' * b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
BMPC_HD __forceinline__ bool mat3_inv(const double* a, double* r) {
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
BMPC_HD __forceinline__ void eul2rotm(const double* e, double* R) {
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
// leg Jacobian (MPC.py:306-365) and low-level torque map (MPC.py:426-470) for one leg
// ------------------------------------------------------------------------------------
BMPC_HD __forceinline__ void leg_jacobian(const double* q, double side, double* J /*6x5 row-major*/) {
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
BMPC_HD __forceinline__ void lowlevel_leg_inl(const DevParams& p, const double* x_fb, double t, const double* pf_w,
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

// out-of-line copy for the kernels where code size matters more than the call (taking the address of the kernel
// parameter block, however, turns every later parameter read into a generic load: the lane kernels use the inline form)
BMPC_HD inline __noinline__ void lowlevel_leg(const DevParams& p, const double* x_fb, double t, const double* pf_w,
                                    const double* q, const double* qd, const double* R, int leg, double c_leg,
                                    const double* u, double* tau_leg) {
    lowlevel_leg_inl(p, x_fb, t, pf_w, q, qd, R, leg, c_leg, u, tau_leg);
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
