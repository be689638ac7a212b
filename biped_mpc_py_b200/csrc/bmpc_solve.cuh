// The fused per-instance MPC kernel (assembly + interior-point solve + torque map).
#pragma once
#include "bmpc_kernels.cuh"

namespace bmpc {

// work_list[bucket][...] holds instance indices; work_count[bucket] how many.
template <int HZ, int SMAX, int LB, int NT>
__global__ void __launch_bounds__(NT) mpc_tick_kernel(const __grid_constant__ DevParams p, const IoPtrs io,
                                                      const int* __restrict__ work_list,
                                                      const int* __restrict__ work_count) {
    using L = Layout<HZ, SMAX, LB>;
    constexpr int RPT = (MAXROWS * SMAX + NT - 1) / NT;  // inequality rows per thread
    static_assert(L::N + 1 <= NT, "one reduced variable per thread");
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x;

    double* s_in = sm + L::o_in;
    double* cur = sm + L::o_cur;
    double* xref = sm + L::o_xref;
    double* rinv = sm + L::o_rinv;
    double* psum = sm + L::o_psum;
    double* iwinv = sm + L::o_iwinv;
    double* err = sm + L::o_err;
    double* footv = sm + L::o_footv;
    double* rotn = sm + L::o_rot;
    double* Wm = sm + L::o_W;
    double* Wp = sm + L::o_Wp;
    double* Vp = sm + L::o_Vp;
    double* Cb = sm + L::o_Cb;
    double* rb = sm + L::o_rb;
    double* ub = sm + L::o_ub;
    double* Hp = sm + L::o_H;
    double* Mp = sm + L::o_M;
    double* gv = sm + L::o_g;
    double* uv = sm + L::o_u;
    double* duv = sm + L::o_du;
    double* xv = sm + L::o_x;
    double* tv = sm + L::o_t;
    double* inv = sm + L::o_inv;
    double* wrow = sm + L::o_wrow;
    double* drow = sm + L::o_drow;
    double* red = sm + L::o_red;
    int* blk_stage = reinterpret_cast<int*>(sm + L::o_int);
    int* blk_foot = blk_stage + SMAX;
    int* blockOf = blk_foot + SMAX;   // [HZ][2]
    int* footsel = blockOf + 2 * HZ;  // [HZ]
    int* cont = footsel + HZ;         // [HZ][2]
    int* misc = cont + 2 * HZ;        // [0]=S
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::o_bar);

    const int count = *work_count;
    if ((int)blockIdx.x >= count) return;

    const int mb = p.mb;
    const double dt = p.dt;

    // ---- TMA staging of the per-instance inputs -------------------------------------
    auto issue_loads = [&](int inst, int buf) {
        // one elected thread arms the barrier and issues 5 bulk copies (or 3 without low-level inputs)
        double* dst = s_in + buf * IN_DOUBLES;
        uint32_t bytes = 96 + 48 + (io.do_lowlevel ? (80 + 80 + 48) : 0);
        mbar_expect_tx(&bars[buf], bytes);
        tma_load_1d(dst, io.x_fb + (size_t)inst * 12, 96, &bars[buf]);
        tma_load_1d(dst + 12, io.foot + (size_t)inst * 6, 48, &bars[buf]);
        if (io.do_lowlevel) {
            tma_load_1d(dst + 18, io.q + (size_t)inst * 10, 80, &bars[buf]);
            tma_load_1d(dst + 28, io.qd + (size_t)inst * 10, 80, &bars[buf]);
            tma_load_1d(dst + 38, io.pf_w + (size_t)inst * 6, 48, &bars[buf]);
        }
    };
    if (io.use_tma) {
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) issue_loads(work_list[blockIdx.x], 0);
    }
    uint32_t parity[2] = {0u, 0u};
    int buf = 0;

    for (int w = blockIdx.x; w < count; w += gridDim.x, buf ^= 1) {
        const int inst = work_list[w];
        // ---- 0. inputs -> cur[] ------------------------------------------------------
        if (io.use_tma) {
            mbar_wait(&bars[buf], parity[buf]);
            parity[buf] ^= 1u;
            if (tid < IN_DOUBLES) cur[tid] = s_in[buf * IN_DOUBLES + tid];
        } else {
            if (tid < 12) cur[tid] = io.x_fb[(size_t)inst * 12 + tid];
            else if (tid < 18) cur[tid] = io.foot[(size_t)inst * 6 + tid - 12];
            else if (io.do_lowlevel && tid < 28) cur[tid] = io.q[(size_t)inst * 10 + tid - 18];
            else if (io.do_lowlevel && tid < 38) cur[tid] = io.qd[(size_t)inst * 10 + tid - 28];
            else if (io.do_lowlevel && tid < 44) cur[tid] = io.pf_w[(size_t)inst * 6 + tid - 38];
        }
        if (tid >= 64 && tid < 64 + 2 * HZ) cont[tid - 64] = io.contact[(size_t)inst * 2 * HZ + tid - 64] ? 1 : 0;
        __syncthreads();
        // prefetch the next instance into the other buffer while this one is solved
        if (io.use_tma && tid == 0 && w + (int)gridDim.x < count) issue_loads(work_list[w + gridDim.x], buf ^ 1);

        const double* x_fb = cur;
        const double* foot = cur + 12;
        const int phase_k = io.phase_k[inst];

        // ---- 1. references and per-stage dynamics pieces (MPC.py:61-109, 148-185) ----
        bool bad = false;
        if (tid < 12) bad = !isfinite(x_fb[tid]);
        if (tid >= 12 && tid < 18) bad = !isfinite(foot[tid - 12]);
        if (tid == 0) {
            // block list: stance foot-stages in (stage, foot) order
            int S = 0;
            for (int s = 0; s < HZ; ++s)
                for (int l = 0; l < 2; ++l) {
                    int b = -1;
                    if (cont[2 * s + l]) {
                        if (S < SMAX) {
                            blk_stage[S] = s;
                            blk_foot[S] = l;
                            b = S;
                        }
                        ++S;
                    }
                    blockOf[2 * s + l] = b;
                }
            misc[0] = S;
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
            eul2rotm(x_fb, rotn);
        }
        if (tid >= 32 && tid < 32 + HZ) {
            const int k = tid - 32;
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
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        const int S = misc[0];
        const int n = S * LB;
        const int m = S * mb;
        if (any_bad || S > SMAX) {
            // refuse: zero outputs, flag
            for (int i = tid; i < HZ * 12; i += NT) io.controls[(size_t)inst * HZ * 12 + i] = 0.0;
            if (io.states)
                for (int i = tid; i < HZ * 13; i += NT) io.states[(size_t)inst * HZ * 13 + i] = 0.0;
            if (io.tau && tid < 10) io.tau[(size_t)inst * 10 + tid] = 0.0;
            if (io.fric_active && tid < HZ) io.fric_active[(size_t)inst * HZ + tid] = 0;
            if (tid == 0) {
                io.status[inst] = 3;
                io.iters[inst] = 0;
                if (io.resid) io.resid[2 * inst] = 0.0, io.resid[2 * inst + 1] = 0.0;
            }
            __syncthreads();
            continue;
        }

        // prefix sums P_k = sum_{l=1..k} Rinv_l and the per-block input maps
        if (tid < HZ) {
            double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int l = 1; l <= tid; ++l)
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] += rinv[9 * l + i];
#pragma unroll
            for (int i = 0; i < 9; ++i) psum[9 * tid + i] = acc[i];
        }
        if (tid >= 32 && tid < 32 + S) {
            const int j = tid - 32, s = blk_stage[j], l = blk_foot[j];
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
        if (tid >= 64 && tid < 64 + mb) {
            const int k = tid - 64;
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
        // strictly feasible start inside one block's polytope
        if (tid == 96) {
            double u6[6];
            for (int c = 0; c < 6; ++c) {
                const double lo = p.lo6[c], hi = p.hi6[c];
                double v0 = fmin(fmax(0.0, lo + 0.1 * (hi - lo)), hi - 0.1 * (hi - lo));
                u6[c] = (hi > lo) ? v0 : lo;
            }
            u6[2] = p.lo6[2] + p.init_fz_frac * (p.hi6[2] - p.lo6[2]);
            for (int c = 0; c < 2; ++c) {
                const double lo = fmax(p.lo6[c], -p.mu * u6[2]), hi = fmin(p.hi6[c], p.mu * u6[2]);
                if (p.hi6[c] > p.lo6[c]) u6[c] = 0.5 * (lo + hi);
            }
#pragma unroll
            for (int c = 0; c < LB; ++c) ub[c] = u6[p.comps[c]];
        }
        __syncthreads();

        // free response error e_i = X_i(u=0, pinned at bound) - x_ref_i
        if (tid < HZ) {
            const int i = tid;
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
        __syncthreads();

        // ---- 2. condensed Hessian blocks and gradient ---------------------------------
        // block pair (jr >= jc): sum_i Gr' Qth Gc + cpp V'QpV + cnt (W'QwW + V'QvV) (+R)
        {
            const int npairs = S * (S + 1) / 2;
            const double vm = dt / p.mass;
            for (int pi = tid; pi < npairs; pi += NT) {
                int jr = (int)((sqrtf(8.0f * (float)pi + 1.0f) - 1.0f) * 0.5f);
                while (jr * (jr + 1) / 2 > pi) --jr;
                while ((jr + 1) * (jr + 2) / 2 <= pi) ++jr;
                const int jc = pi - jr * (jr + 1) / 2;
                const int sr = blk_stage[jr], sc = blk_stage[jc];  // sc <= sr
                const double* Wr = Wm + jr * 3 * LB;
                const double* Wc = Wm + jc * 3 * LB;
                double blk[LB][LB];
#pragma unroll
                for (int a = 0; a < LB; ++a)
#pragma unroll
                    for (int b = 0; b < LB; ++b) blk[a][b] = 0.0;
                double cpp = 0.0;
                for (int i = sr; i < HZ; ++i) {
                    cpp += (double)(i - sr) * (double)(i - sc);
                    if (i == sr) continue;  // P_i - P_sr = 0: no orientation coupling yet
                    const double* Pi = psum + 9 * i;
                    const double* Pr = psum + 9 * sr;
                    const double* Pc = psum + 9 * sc;
                    double Dr[9], Dc[9];
#pragma unroll
                    for (int x = 0; x < 9; ++x) Dr[x] = dt * (Pi[x] - Pr[x]), Dc[x] = dt * (Pi[x] - Pc[x]);
                    double Gr[3][LB], Gc[3][LB];
#pragma unroll
                    for (int x = 0; x < 3; ++x)
#pragma unroll
                        for (int a = 0; a < LB; ++a) {
                            Gr[x][a] = Dr[3 * x] * Wr[a] + Dr[3 * x + 1] * Wr[LB + a] + Dr[3 * x + 2] * Wr[2 * LB + a];
                            Gc[x][a] = Dc[3 * x] * Wc[a] + Dc[3 * x + 1] * Wc[LB + a] + Dc[3 * x + 2] * Wc[2 * LB + a];
                        }
#pragma unroll
                    for (int x = 0; x < 3; ++x)
#pragma unroll
                        for (int a = 0; a < LB; ++a) {
                            const double qa = p.Q[x] * Gr[x][a];
#pragma unroll
                            for (int b = 0; b < LB; ++b) blk[a][b] += qa * Gc[x][b];
                        }
                }
                const double cnt = (double)(HZ - sr);
#pragma unroll
                for (int a = 0; a < LB; ++a)
#pragma unroll
                    for (int b = 0; b < LB; ++b) {
                        double acc = 0.0;
#pragma unroll
                        for (int x = 0; x < 3; ++x) acc += p.Q[6 + x] * Wr[x * LB + a] * Wc[x * LB + b];
                        blk[a][b] += cnt * acc;
                    }
                // force components: V = (dt/m) e_a  -> diagonal terms only
#pragma unroll
                for (int a = 0; a < LB; ++a) {
                    const int ca = p.comps[a];
                    if (ca < 3) {
#pragma unroll
                        for (int b = 0; b < LB; ++b)
                            if (p.comps[b] == ca)
                                blk[a][b] += vm * vm * (dt * dt * cpp * p.Q[3 + ca] + cnt * p.Q[9 + ca]);
                    }
                }
                if (jr == jc) {
                    const int l = blk_foot[jr];
#pragma unroll
                    for (int a = 0; a < LB; ++a) {
                        const int ca = p.comps[a];
                        blk[a][a] += p.R[(ca < 3) ? (3 * l + ca) : (6 + 3 * l + ca - 3)];
                    }
                }
#pragma unroll
                for (int a = 0; a < LB; ++a)
#pragma unroll
                    for (int b = 0; b < LB; ++b)
                        if (jr != jc || b <= a) Hp[tri(jr * LB + a, jc * LB + b)] = blk[a][b];
            }
            // gradient: thread per block (threads from the top so they overlap with the pair loop)
            if (tid >= NT - S) {
                const int j = NT - 1 - tid, s = blk_stage[j];
                const double* Wj = Wm + j * 3 * LB;
                double gj[LB];
#pragma unroll
                for (int a = 0; a < LB; ++a) gj[a] = 0.0;
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
                    // (P_i - P_s)' (Qth e_th)
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
                    gj[a] = acc;
                }
#pragma unroll
                for (int a = 0; a < LB; ++a) gv[j * LB + a] = gj[a];
            }
        }
        __syncthreads();

        if (io.dbg_H != nullptr && w == 0) {
            const int nmax = 12 * HZ;
            for (int e = tid; e < n * n; e += NT) {
                const int i = e / n, j = e - i * n;
                io.dbg_H[i * nmax + j] = (j <= i) ? Hp[tri(i, j)] : Hp[tri(j, i)];
            }
            if (tid < n) io.dbg_g[tid] = gv[tid];
            if (tid == 0) io.dbg_n[0] = n;
        }

        // ---- 3. interior point (Mehrotra predictor-corrector), FP64 ---------------------
        // rows owned by this thread stay in registers across phases
        int rj[RPT], rk[RPT];
        double rs_[RPT], rl_[RPT];  // slack, multiplier
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
            const int r = tid + rr * NT;
            rj[rr] = (r < m) ? r / mb : -1;
            rk[rr] = (r < m) ? r - (r / mb) * mb : 0;
        }
        auto row_dot = [&](int rr, const double* v) {
            const double* cb = Cb + rk[rr] * LB;
            const double* vv = v + rj[rr] * LB;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < LB; ++c) acc += cb[c] * vv[c];
            return acc;
        };
        auto col_gather = [&](int i, const double* wr) {  // (C' w)_i for variable i
            const int j = i / LB, c = i - j * LB;
            const double* ww = wr + j * mb;
            double acc = 0.0;
            for (int k = 0; k < mb; ++k) acc += Cb[k * LB + c] * ww[k];
            return acc;
        };

        if (tid < n) uv[tid] = ub[tid % LB];
        double mu0_part = 0.0;
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
            rs_[rr] = 1.0, rl_[rr] = 1.0;
            if (rj[rr] >= 0) {
                const double* cb = Cb + rk[rr] * LB;
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < LB; ++c) acc += cb[c] * ub[c];
                double sl = rb[rk[rr]] - acc;
                if (!(sl > 1e-3)) sl = 1.0;  // infeasible start for this row: handled through rp
                rs_[rr] = sl;
                mu0_part += sl;
            }
        }
        int status = 1, it = 0;
        double mu = 0.0, rdmax = 0.0;
        if (n > 0) {
            const double mu0 = block_sum<NT>(mu0_part, red) / (double)m;
#pragma unroll
            for (int rr = 0; rr < RPT; ++rr) rl_[rr] = mu0 / rs_[rr];

            for (it = 1; it <= p.max_iter; ++it) {
                // t = Hc u  (packed symmetric matvec, one row per thread)
                if (tid < n) {
                    double acc = 0.0;
                    const double* row = Hp + tri(tid, 0);
                    for (int j = 0; j <= tid; ++j) acc += row[j] * uv[j];
                    for (int j = tid + 1; j < n; ++j) acc += Hp[tri(j, tid)] * uv[j];
                    tv[tid] = acc;
                }
                double rp_[RPT], d_[RPT], is_[RPT];
                double part = 0.0;
#pragma unroll
                for (int rr = 0; rr < RPT; ++rr) {
                    rp_[rr] = 0.0, d_[rr] = 0.0, is_[rr] = 1.0;
                    if (rj[rr] >= 0) {
                        const int r = tid + rr * NT;
                        is_[rr] = 1.0 / rs_[rr];
                        d_[rr] = rl_[rr] * is_[rr];
                        rp_[rr] = row_dot(rr, uv) + rs_[rr] - rb[rk[rr]];
                        wrow[r] = rl_[rr];
                        drow[r] = d_[rr];
                        part += rs_[rr] * rl_[rr];
                    }
                }
                __syncthreads();
                double rd_i = 0.0;
                if (tid < n) rd_i = tv[tid] + gv[tid] + col_gather(tid, wrow);
                mu = block_sum<NT>(part, red) / (double)m;
                rdmax = block_max<NT>(fabs(rd_i), red);
                if (mu <= p.mu_tol && rdmax <= p.rd_tol) {
                    status = 0;
                    break;
                }
                // M = Hc + blockdiag(Cb' diag(d_j) Cb); augmented row = predictor rhs
#pragma unroll
                for (int rr = 0; rr < RPT; ++rr)
                    if (rj[rr] >= 0) wrow[tid + rr * NT] = d_[rr] * rp_[rr] - rl_[rr];
                for (int e = tid; e < L::HP; e += NT)
                    if (e < n * (n + 1) / 2) Mp[e] = Hp[e];
                __syncthreads();
                for (int e = tid; e < S * LB * LB; e += NT) {
                    const int j = e / (LB * LB), ab = e - j * LB * LB, a = ab / LB, b = ab - a * LB;
                    if (b <= a) {
                        const double* dd = drow + j * mb;
                        double acc = 0.0;
                        for (int k = 0; k < mb; ++k) acc += Cb[k * LB + a] * Cb[k * LB + b] * dd[k];
                        Mp[tri(j * LB + a, j * LB + b)] += acc;
                    }
                }
                if (tid < n) Mp[tri(n, tid)] = -rd_i - col_gather(tid, wrow);
                __syncthreads();
                if (!chol_aug<NT>(Mp, n, inv)) {
                    status = 2;
                    break;
                }
                if (tid < n) xv[tid] = Mp[tri(n, tid)];
                __syncthreads();
                solve_backward<NT>(Mp, n, inv, xv);  // xv = du_aff
                double dsa[RPT], dla[RPT];
                double ratio = 0.0;
#pragma unroll
                for (int rr = 0; rr < RPT; ++rr) {
                    dsa[rr] = 0.0, dla[rr] = 0.0;
                    if (rj[rr] >= 0) {
                        dsa[rr] = -rp_[rr] - row_dot(rr, xv);
                        dla[rr] = -rl_[rr] - d_[rr] * dsa[rr];
                        ratio = fmax(ratio, fmax(-dsa[rr] * is_[rr], -dla[rr] / rl_[rr]));
                    }
                }
                if (tid < n) duv[tid] = xv[tid];
                ratio = block_max<NT>(ratio, red);
                const double a_aff = (ratio > 1.0) ? 1.0 / ratio : 1.0;
                part = 0.0;
#pragma unroll
                for (int rr = 0; rr < RPT; ++rr)
                    if (rj[rr] >= 0) part += (rs_[rr] + a_aff * dsa[rr]) * (rl_[rr] + a_aff * dla[rr]);
                const double mu_aff = block_sum<NT>(part, red) / (double)m;
                double sigma = mu_aff / mu;
                sigma = sigma * sigma * sigma;
                // corrector: M dcorr = +C' ((dsa*dla - sigma*mu) / s)
                double wc_[RPT];
#pragma unroll
                for (int rr = 0; rr < RPT; ++rr) {
                    wc_[rr] = 0.0;
                    if (rj[rr] >= 0) {
                        wc_[rr] = (dsa[rr] * dla[rr] - sigma * mu) * is_[rr];
                        wrow[tid + rr * NT] = wc_[rr];
                    }
                }
                __syncthreads();
                if (tid < n) xv[tid] = col_gather(tid, wrow);
                __syncthreads();
                solve_forward<NT>(Mp, n, inv, xv);
                solve_backward<NT>(Mp, n, inv, xv);
                if (tid < n) duv[tid] += xv[tid];
                __syncthreads();
                double ds_[RPT], dl_[RPT];
                ratio = 0.0;
#pragma unroll
                for (int rr = 0; rr < RPT; ++rr) {
                    ds_[rr] = 0.0, dl_[rr] = 0.0;
                    if (rj[rr] >= 0) {
                        ds_[rr] = -rp_[rr] - row_dot(rr, duv);
                        dl_[rr] = -rl_[rr] - wc_[rr] - d_[rr] * ds_[rr];
                        ratio = fmax(ratio, fmax(-ds_[rr] * is_[rr], -dl_[rr] / rl_[rr]));
                    }
                }
                ratio = block_max<NT>(ratio, red);
                const double alpha = (ratio > 0.995) ? 0.995 / ratio : 1.0;
                if (!isfinite(alpha) || !isfinite(ratio)) {
                    status = 2;
                    break;
                }
                if (tid < n) uv[tid] += alpha * duv[tid];
#pragma unroll
                for (int rr = 0; rr < RPT; ++rr)
                    if (rj[rr] >= 0) {
                        rs_[rr] += alpha * ds_[rr];
                        rl_[rr] += alpha * dl_[rr];
                    }
                __syncthreads();
            }
            if (it > p.max_iter) it = p.max_iter;
        } else {
            status = 0;
        }
        __syncthreads();

        // ---- 4. outputs ------------------------------------------------------------------
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
            if (s == 0) xv[c12] = val;  // first-stage input for the torque map (xv is free now)
        }
        const double uscale = fmax(1.0, block_max<NT>(umax_part, red));
        // predicted states (h,13): X_i = free response + sum_j dX_i/du_j u_j
        if (io.states && tid < HZ) {
            const int i = tid;
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
        // non-trivially active friction rows per stage
        if (io.fric_active && tid >= 32 && tid < 32 + HZ) {
            const int s = tid - 32;
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
        __syncthreads();
        // joint torques from the first-stage input (MPC.py:444-470), one thread per leg
        if (io.do_lowlevel && io.tau && (tid == 0 || tid == 32)) {
            const int leg = tid >> 5;
            double tl[5];
            lowlevel_leg(p, x_fb, io.t_swing[inst], cur + 38, cur + 18, cur + 28, rotn, leg, (double)cont[leg], xv, tl);
#pragma unroll
            for (int c = 0; c < 5; ++c) io.tau[(size_t)inst * 10 + 5 * leg + c] = tl[c];
        }
        if (tid == 64) {
            io.status[inst] = status;
            io.iters[inst] = it;
            if (io.resid) io.resid[2 * inst] = mu, io.resid[2 * inst + 1] = rdmax;
        }
        __syncthreads();
    }
}

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
