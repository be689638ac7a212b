// Closed-loop batched rollout: gait clock, single-rigid-body plant step and
// touchdown foothold update around the fused MPC tick.  The reference has no loop (MPC.py:475-495
// runs one tick); the rules R1-R6 are stated in oracle/rollout.py and DESIGN.md and are built only
// from the reference's own pieces: the contact table (MPC.py:52-55), the discretised dynamics
// (MPC.py:148-185) and the swing controller's target (MPC.py:427-435).
#pragma once
#include "bmpc_kernels.cuh"

namespace bmpc {

struct RolloutPtrs {
    double* x;            // [N,12] state, advanced in place
    double* foot;         // [N,6] foot positions (= pf_w), updated at touchdown
    int32_t* tick;        // [N] integer gait clock, advanced in place
    const uint8_t* gait;  // [N] 1 = walking, 0 = standing
    const double* q;      // [N,10] joint angles (held fixed; used by the torque map and the fall reset)
    // per-tick solver inputs / outputs (library workspace)
    uint8_t* contact;     // [N,h,2]
    int32_t* phase_k;     // [N]
    double* t_swing;      // [N]
    const double* controls;  // [N,h,12]
    const double* tau;       // [N,10]
    const int32_t* status;   // [N]
    const int32_t* iters;    // [N]
    // logs of the first n_log robots (nullable)
    int n_log;
    double* x_log;        // [ticks+1,n_log,12]
    double* foot_log;     // [ticks+1,n_log,6]
    double* u0_log;       // [ticks,n_log,12]
    double* tau_log;      // [ticks,n_log,10]
    unsigned long long* stats;  // [8]: sum iters, not-optimal, bad input, max iters, robot-ticks, polish-first hits, falls
};

constexpr int GAIT_PERIOD = 10;  // 5 ticks left stance, 5 right (MPC.py:52-55)

__device__ __forceinline__ int gait_contact(int tick, int leg) {  // walking table, MPC.py:52-55
    int k = tick % GAIT_PERIOD;
    if (k < 0) k += GAIT_PERIOD;
    const int left = k < GAIT_PERIOD / 2;
    return leg == 0 ? left : !left;
}

// R1 + R2: contact schedule, phase and swing clock of the coming tick; logs the state it starts from
__global__ void rollout_prepare_kernel(const __grid_constant__ DevParams p, const RolloutPtrs r, int n, int step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int T = r.tick[i];
    const int walking = r.gait[i] != 0;
    int k = T % GAIT_PERIOD;
    if (k < 0) k += GAIT_PERIOD;
    r.phase_k[i] = k;
    r.t_swing[i] = (double)T * p.dt;
    for (int s = 0; s < p.h; ++s)
        for (int l = 0; l < 2; ++l)
            r.contact[((size_t)i * p.h + s) * 2 + l] = walking ? (uint8_t)gait_contact(T + s, l) : (uint8_t)1;
    if (i < r.n_log && r.x_log) {
        for (int c = 0; c < 12; ++c) r.x_log[((size_t)step * r.n_log + i) * 12 + c] = r.x[(size_t)i * 12 + c];
        for (int c = 0; c < 6; ++c) r.foot_log[((size_t)step * r.n_log + i) * 6 + c] = r.foot[(size_t)i * 6 + c];
    }
}

// R5: x+ = A_0 [x;1] + B_0 u_0 with the reference's discretisation at the current state and feet
// (MPC.py:148-185; the dynamics read x[0] as yaw, x[1] pitch, x[2] roll, MPC.py:151-153)
__device__ __forceinline__ void srb_step(const DevParams& p, const double* x, const double* foot, const double* u,
                                         double* xn) {
    double sz, cz, sy, cy, sx, cx;
    sincos(x[0], &sz, &cz);
    sincos(x[1], &sy, &cy);
    sincos(x[2], &sx, &cx);
    double rot[9];  // Rx(roll) Ry(pitch) Rz(yaw): extrinsic 'zyx', MPC.py:156
    rot[0] = cy * cz, rot[1] = -cy * sz, rot[2] = sy;
    rot[3] = sx * sy * cz + cx * sz, rot[4] = -sx * sy * sz + cx * cz, rot[5] = -sx * cy;
    rot[6] = -cx * sy * cz + sx * sz, rot[7] = cx * sy * sz + sx * cz, rot[8] = cx * cy;
    double tmp[9], iw[9], iwi[9];
    mat3_mul(p.inertia, rot, tmp);
    mat3_tmul(rot, tmp, iw);  // Rot' I Rot, MPC.py:157
    mat3_inv(iw, iwi);
    const double icp = 1.0 / cy;
    const double ri[9] = {cz * icp, sz * icp, 0.0, -sz, cz, 0.0, cz * sy * icp, sz * sy * icp, 1.0};  // MPC.py:160-164
    // net moment about the centre of mass: sum_l (foot_l - com) x f_l + m_l   (MPC.py:174-179)
    double mom[3] = {0, 0, 0}, frc[3] = {0, 0, 0};
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const double r0 = foot[3 * l] - x[3], r1 = foot[3 * l + 1] - x[4], r2 = foot[3 * l + 2] - x[5];
        const double f0 = u[3 * l], f1 = u[3 * l + 1], f2 = u[3 * l + 2];
        mom[0] += r1 * f2 - r2 * f1 + u[6 + 3 * l];
        mom[1] += r2 * f0 - r0 * f2 + u[6 + 3 * l + 1];
        mom[2] += r0 * f1 - r1 * f0 + u[6 + 3 * l + 2];
        frc[0] += f0, frc[1] += f1, frc[2] += f2;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        xn[a] = x[a] + p.dt * (ri[3 * a] * x[6] + ri[3 * a + 1] * x[7] + ri[3 * a + 2] * x[8]);
        xn[3 + a] = x[3 + a] + p.dt * x[9 + a];
        xn[6 + a] = x[6 + a] + p.dt * (iwi[3 * a] * mom[0] + iwi[3 * a + 1] * mom[1] + iwi[3 * a + 2] * mom[2]);
        xn[9 + a] = x[9 + a] + p.dt * (frc[a] / p.mass);
    }
    xn[11] -= p.dt * p.g;  // Ac[11,12] = -g acting on the homogeneous 1 (MPC.py:171)
}

// R5 + R6 + clock, one thread per robot; also the logs and the solver statistics of the tick
__global__ void rollout_advance_kernel(const __grid_constant__ DevParams p, const RolloutPtrs r, int n, int step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long it = 0ull, notopt = 0ull, bad = 0ull, hit = 0ull, falls = 0ull;
    unsigned itmax = 0u;
    if (i < n) {
        double x[12], xn[12], foot[6], u[12];
        for (int c = 0; c < 12; ++c) x[c] = r.x[(size_t)i * 12 + c], u[c] = r.controls[(size_t)i * p.h * 12 + c];
        for (int c = 0; c < 6; ++c) foot[c] = r.foot[(size_t)i * 6 + c];
        srb_step(p, x, foot, u, xn);
        const int T = r.tick[i];
        // R7: fallen (or overflowed) robots go back to the reference's initial state (MPC.py:13) with
        // the feet of their own joint angles (MPC.py:406-424 at zero orientation: R = I)
        bool fall = false;
#pragma unroll
        for (int c = 0; c < 12; ++c) fall = fall || !isfinite(xn[c]);
        fall = fall || fabs(xn[0]) > 0.8 || fabs(xn[1]) > 0.8 || fabs(xn[2]) > 0.8 || xn[5] < 0.25 || xn[5] > 1.0;
        if (fall) {
#pragma unroll
            for (int c = 0; c < 12; ++c) xn[c] = (c == 5) ? 0.53 : 0.0;
            double e0[3] = {0.0, 0.0, 0.0}, R[9];
            eul2rotm(e0, R);
            for (int l = 0; l < 2; ++l) {
                const double side = (l == 0) ? 1.0 : -1.0;
                double qq[5], pb[3];
                for (int c = 0; c < 5; ++c) qq[c] = r.q[(size_t)i * 10 + 5 * l + c];
                foot_body(qq, side, pb);
                pb[0] += p.hip[0], pb[1] += side * p.hip[1], pb[2] += p.hip[2];
                for (int a = 0; a < 3; ++a) foot[3 * l + a] = xn[3 + a] + R[a] * pb[0] + R[3 + a] * pb[1] + R[6 + a] * pb[2];
            }
            falls = 1ull;
        } else if (r.gait[i] != 0) {
            const double hh = (double)p.h;
#pragma unroll
            for (int l = 0; l < 2; ++l)
                if (gait_contact(T + 1, l) && !gait_contact(T, l)) {
                    const double side = (l == 0) ? 1.0 : -1.0;
                    foot[3 * l] = xn[3] + xn[9] * 1 / 2 * hh / 2 * p.dt + p.kv * (xn[3] - p.x_cmd[3]);
                    foot[3 * l + 1] = xn[4] + xn[10] * 1 / 2 * hh / 2 * p.dt + p.kv * (xn[4] - p.x_cmd[4]) + 0.04 * side;
                    foot[3 * l + 2] = 0.0;
                }
        }
        for (int c = 0; c < 12; ++c) r.x[(size_t)i * 12 + c] = xn[c];
        for (int c = 0; c < 6; ++c) r.foot[(size_t)i * 6 + c] = foot[c];
        r.tick[i] = T + 1;
        if (i < r.n_log && r.u0_log) {
            for (int c = 0; c < 12; ++c) r.u0_log[((size_t)step * r.n_log + i) * 12 + c] = u[c];
            for (int c = 0; c < 10; ++c) r.tau_log[((size_t)step * r.n_log + i) * 10 + c] = r.tau[(size_t)i * 10 + c];
            // the state after the last tick closes the log
            for (int c = 0; c < 12; ++c) r.x_log[((size_t)(step + 1) * r.n_log + i) * 12 + c] = xn[c];
            for (int c = 0; c < 6; ++c) r.foot_log[((size_t)(step + 1) * r.n_log + i) * 6 + c] = foot[c];
        }
        const int st = r.status[i], its = r.iters[i];
        it = (unsigned long long)its;
        itmax = (unsigned)its;
        notopt = st != 0;
        bad = st == 3;
        hit = (st == 0 && its == 0);
    }
    // warp reduction, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        it += __shfl_xor_sync(0xffffffffu, it, o);
        notopt += __shfl_xor_sync(0xffffffffu, notopt, o);
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        hit += __shfl_xor_sync(0xffffffffu, hit, o);
        falls += __shfl_xor_sync(0xffffffffu, falls, o);
        itmax = max(itmax, __shfl_xor_sync(0xffffffffu, itmax, o));
    }
    if ((threadIdx.x & 31) == 0 && r.stats) {
        const int cnt = min(32, n - (i & ~31));
        if (cnt > 0) {
            atomicAdd(&r.stats[0], it);
            atomicAdd(&r.stats[1], notopt);
            atomicAdd(&r.stats[2], bad);
            atomicMax(&r.stats[3], (unsigned long long)itmax);
            atomicAdd(&r.stats[4], (unsigned long long)cnt);
            atomicAdd(&r.stats[5], hit);
            atomicAdd(&r.stats[6], falls);
        }
    }
}

}  // namespace bmpc
