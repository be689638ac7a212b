// Small kernels: instance bucketing, stand-alone torque map, batched forward kinematics, FMA peak.
#pragma once
#include "bmpc_kernels.cuh"

namespace bmpc {

// ------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------

// bucket instances by number of stance foot-stages: list0 = S <= h, list1 = the rest
__global__ void classify_kernel(const uint8_t* __restrict__ contact, int n, int h, int list_stride,
                                int* __restrict__ lists, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int S = 0;
    for (int k = 0; k < 2 * h; ++k) S += contact[(size_t)i * 2 * h + k] ? 1 : 0;
    const int b = (S <= h) ? 0 : 1;
    // order inside a bucket does not affect any result (instances are independent)
    const int slot = atomicAdd(&counts[b], 1);
    lists[(size_t)b * list_stride + slot] = i;
}

// instances of a work list that were not certified optimal (and are not bad input): they are re-solved by the
// dense fallback kernel (the stage-wise backend loses accuracy earlier when barrier weights exceed ~1e10)
__global__ void collect_uncertified_kernel(const int* __restrict__ list, const int* __restrict__ count,
                                           const int32_t* __restrict__ status, int* __restrict__ out_list,
                                           int* __restrict__ out_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const int inst = list[i];
    const int st = status[inst];
    if (st == 1 || st == 2) out_list[atomicAdd(out_count, 1)] = inst;
}

// after a lane-per-robot front end: the robots it did not certify, or the whole list when the class was below the front end's
// minimum size (it did not run then and status[] is stale)
__global__ void collect_or_all_kernel(const int* __restrict__ list, const int* __restrict__ count, const int32_t* __restrict__ status,
                                      int* __restrict__ out_list, int* __restrict__ out_count, int min_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = *count;
    if (i >= c) return;
    const int inst = list[i];
    if (c < min_count) {
        out_list[i] = inst;
        if (i == 0) *out_count = c;
        return;
    }
    const int st = status[inst];
    if (st == 1 || st == 2) out_list[atomicAdd(out_count, 1)] = inst;
}

// lowLevelControl only (MPC.py:444-470): one thread per (instance, leg)
__global__ void lowlevel_kernel(const __grid_constant__ DevParams p, int n, const double* __restrict__ x_fb,
                                const double* __restrict__ t_swing, const double* __restrict__ pf_w,
                                const double* __restrict__ q, const double* __restrict__ qd,
                                const uint8_t* __restrict__ contact0, const double* __restrict__ u0,
                                double* __restrict__ tau) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gid >> 1, leg = gid & 1;
    if (i >= n) return;
    double xf[12], pf[6], qq[10], qv[10], u[12], R[9], tl[5];
    for (int k = 0; k < 12; ++k) xf[k] = x_fb[(size_t)i * 12 + k], u[k] = u0[(size_t)i * 12 + k];
    for (int k = 0; k < 6; ++k) pf[k] = pf_w[(size_t)i * 6 + k];
    for (int k = 0; k < 10; ++k) qq[k] = q[(size_t)i * 10 + k], qv[k] = qd[(size_t)i * 10 + k];
    eul2rotm(xf, R);
    lowlevel_leg(p, xf, t_swing[i], pf, qq, qv, R, leg, contact0[(size_t)i * 2 + leg] ? 1.0 : 0.0, u, tl);
    for (int c = 0; c < 5; ++c) tau[(size_t)i * 10 + 5 * leg + c] = tl[c];
}

// getFootPositionWorld (MPC.py:406-424): one thread per (instance, leg)
__global__ void foot_positions_kernel(const __grid_constant__ DevParams p, int n, const double* __restrict__ x_fb,
                                      const double* __restrict__ q, double* __restrict__ pf_w) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gid >> 1, leg = gid & 1;
    if (i >= n) return;
    const double side = leg == 0 ? 1.0 : -1.0;
    double e[3] = {x_fb[(size_t)i * 12], x_fb[(size_t)i * 12 + 1], x_fb[(size_t)i * 12 + 2]};
    double R[9], qq[5], pb[3];
    eul2rotm(e, R);
    for (int k = 0; k < 5; ++k) qq[k] = q[(size_t)i * 10 + 5 * leg + k];
    foot_body(qq, side, pb);
    pb[0] += p.hip[0];
    pb[1] += side * p.hip[1];
    pb[2] += p.hip[2];
    for (int a = 0; a < 3; ++a)  // p_c + R.T @ (pf_b + hip)  (MPC.py:423)
        pf_w[(size_t)i * 6 + 3 * leg + a] = x_fb[(size_t)i * 12 + 3 + a] + R[a] * pb[0] + R[3 + a] * pb[1] + R[6 + a] * pb[2];
}

// register-resident FMA chains on every SM: CUDA-core peak (roofline denominator)
template <typename T>
__global__ void fma_peak_kernel(T* out, int iters) {
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
    T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T b = (T)0.999999, c = (T)1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * b + c, a1 = a1 * b + c, a2 = a2 * b + c, a3 = a3 * b + c;
        a4 = a4 * b + c, a5 = a5 * b + c, a6 = a6 * b + c, a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace bmpc
