// Lane-per-robot MPC tick: ONE THREAD solves one robot, 32 robots per warp execute the same instruction stream.
//
// Why: the warp-per-robot kernel (bmpc_tick.cuh) is bound by instruction delivery — every warp walks a 40 KB loop at
// its own position for one 50-variable problem, ~12 issued warp instructions per useful FP64 FMA lane-op
// (profiles/r1_summary.md).  Here an instruction serves 32 robots, there is no cross-lane communication and no
// barrier, and the linear algebra is the stage-wise (Riccati) form of the same QP: the contact-reduced problem is an
// LQR problem  X_i = A_i X_{i-1} + B_i u_i + c_i  (MPC.py:165-184, 206-214) with block-diagonal input weights, so a
// backward sweep over the 12x12 cost-to-go factors  M = Hc + blockdiag(Cb' D Cb)  in O(h) and Hc is never formed.
// Same algorithm as the tick kernel otherwise (Mehrotra + Gondzio interior point to a loose target, active-set polish,
// KKT certificate; the per-block null-space / multiplier checks are the SAME functions, bmpc_polish.cuh), written as
// plain scalar C++ so that the identical source is unit-tested on the CPU (tests/lane_host.cu) against the oracle.
//
// Per-robot work arrays live in a global workspace interleaved by lane (element i of lane l at ws[i*32 + l]: every
// access of a warp is one fully used 256-byte line pair); on the host the stride is 1.
// A robot this path does not certify keeps status 1 and is re-solved by the warp-per-robot kernels (bmpc.cu).
#pragma once
#include "bmpc_kernels.cuh"
#include "bmpc_polish.cuh"

namespace bmpc {

#ifdef __CUDA_ARCH__
#define BMPC_LS 32
#else
#define BMPC_LS 1
#endif

struct SV {  // strided view of one lane's slice of the workspace
    double* p;
    BMPC_HD __forceinline__ double& operator[](int i) const { return p[(size_t)i * BMPC_LS]; }
    BMPC_HD __forceinline__ SV operator+(int o) const { return SV{p + (size_t)o * BMPC_LS}; }
    // streaming read: data that will not be touched again before it is evicted anyway (the stored factor in the solves)
    BMPC_HD __forceinline__ void sts(int i, double v) const {
#ifdef __CUDA_ARCH__
        __stcs(p + (size_t)i * BMPC_LS, v);
#else
        p[(size_t)i * BMPC_LS] = v;
#endif
    }
    BMPC_HD __forceinline__ double lds(int i) const {
#ifdef __CUDA_ARCH__
        return __ldcs(p + (size_t)i * BMPC_LS);
#else
        return p[(size_t)i * BMPC_LS];
#endif
    }
};

template <int HZ, int NF, int LB>
struct LaneL {
    static constexpr int NU = NF * LB, S = HZ * NF, N = NU * HZ, MBM = 12, M = S * MBM, E = LB * LB;
    static constexpr int o_u = 0, o_du = o_u + N, o_x = o_du + N, o_up = o_x + N, o_rd = o_up + N, o_pp = o_rd + N,
                         o_tv = o_pp + N, o_hd = o_tv + N;
    static constexpr int o_rs = o_hd + N, o_rl = o_rs + M, o_rdw = o_rl + M, o_rp = o_rdw + M, o_rc = o_rp + M,
                         o_rw = o_rc + M;
    static constexpr int o_P = o_rw + M;                 // 12 x 12 cost-to-go
    static constexpr int o_PB = o_P + 144;               // 12 x NU
    static constexpr int o_F = o_PB + 12 * NU;           // NU x 12
    static constexpr int o_G = o_F + 12 * NU;            // NU x NU
    static constexpr int o_K = o_G + NU * NU;            // per stage NU x 12 feedback gains
    static constexpr int o_Lc = o_K + HZ * NU * 12;      // per stage Cholesky factor of G (reciprocal diagonal)
    static constexpr int o_Rt = o_Lc + HZ * NU * NU;     // per stage input weights
    static constexpr int o_Bm = o_Rt + HZ * NU * NU;     // per stage 6 x NU input maps in use (B, or B N in the polish)
    static constexpr int o_B0 = o_Bm + HZ * 6 * NU;      // per stage 6 x NU input maps of the problem
    static constexpr int o_c = o_B0 + HZ * 6 * NU;       // per stage affine term (omega, v rows)
    static constexpr int o_rinv = o_c + HZ * 6;
    static constexpr int o_xref = o_rinv + HZ * 9;
    static constexpr int o_E = o_xref + HZ * 12;         // Q (X_i - xref_i)
    static constexpr int o_Nn = o_E + HZ * 12;           // polish: null-space blocks
    static constexpr int total = o_Nn + S * E;
};

template <int HZ, int NF, int LB>
struct LaneSolver {
    using L = LaneL<HZ, NF, LB>;
    static_assert(LB == 5, "the input-map / weight layouts below are the block layouts of the register-blocked LB = 5 sweep");
    static constexpr int NU = L::NU, S = L::S, N = L::N, E = L::E;
    const DevParams& p;
    SV ws;
    double x_fb[12];
    double Cb[L::MBM * LB], rb[L::MBM], Rd[2][LB];
    int fo[S];  // foot of every block (stage-major)
    // Gondzio's centrality corrector is a per-robot branch: in a warp of 32 robots some lane takes it in nearly every
    // iteration and the other lanes idle through its extra solve.  Measured (262,144 robots, walking class): always 110.9 ms,
    // step < 0.95 (the warp-per-robot setting) 107.5 ms, never 95.9 ms (9.5 instead of 8.6 iterations, each cheaper).
    static constexpr bool kGondzio = false;
    int mb, m;
    double dt;

    BMPC_HD LaneSolver(const DevParams& pp, SV w) : p(pp), ws(w) {}

    // ---- objective gradient  Hc u + g  by a rollout and an adjoint sweep; optionally writes the states -------------
    BMPC_HD void grad(SV u, SV out, double* states) {
        double z[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) z[a] = x_fb[a];
#pragma unroll 1
        for (int i = 0; i < HZ; ++i) {
            SV ri = ws + (L::o_rinv + 9 * i), B = ws + (L::o_B0 + 6 * NU * i), c = ws + (L::o_c + 6 * i), ui = u + NU * i;
            double acc[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[k] = c[k];
#pragma unroll 1
            for (int col = 0; col < NU; ++col) {
                const double uc = ui[col];
#pragma unroll
                for (int k = 0; k < 6; ++k) acc[k] += B[k * NU + col] * uc;
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                z[a] += dt * (ri[3 * a] * z[6] + ri[3 * a + 1] * z[7] + ri[3 * a + 2] * z[8]);
                z[3 + a] += dt * z[9 + a];
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) z[6 + k] += acc[k];
            SV Ei = ws + (L::o_E + 12 * i), xr = ws + (L::o_xref + 12 * i);
#pragma unroll
            for (int a = 0; a < 12; ++a) Ei[a] = p.Q[a] * (z[a] - xr[a]);
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
            SV ri = ws + (L::o_rinv + 9 * i), B = ws + (L::o_B0 + 6 * NU * i), Ei = ws + (L::o_E + 12 * i), ui = u + NU * i,
               oi = out + NU * i;
#pragma unroll
            for (int a = 0; a < 12; ++a) lam[a] += Ei[a];
#pragma unroll 1
            for (int col = 0; col < NU; ++col) {
                double acc = Rd[fo[i * NF + col / LB]][col % LB] * ui[col];
#pragma unroll
                for (int k = 0; k < 6; ++k) acc += B[k * NU + col] * lam[6 + k];
                oi[col] = acc;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                lam[6 + k] += dt * (ri[k] * lam[0] + ri[3 + k] * lam[1] + ri[6 + k] * lam[2]);
                lam[9 + k] += dt * lam[3 + k];
            }
        }
    }

    // P <- A_i' P A_i  (A = I + dt E: a column operation, then a row operation)
    BMPC_HD void congruence(SV P, SV ri) {
        double r9[9];
#pragma unroll
        for (int a = 0; a < 9; ++a) r9[a] = dt * ri[a];
#pragma unroll 1
        for (int r = 0; r < 12; ++r) {
            const double p0 = P[r * 12], p1 = P[r * 12 + 1], p2 = P[r * 12 + 2];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                P[r * 12 + 6 + k] += p0 * r9[k] + p1 * r9[3 + k] + p2 * r9[6 + k];
                P[r * 12 + 9 + k] += dt * P[r * 12 + 3 + k];
            }
        }
#pragma unroll 1
        for (int c = 0; c < 12; ++c) {
            const double p0 = P[c], p1 = P[12 + c], p2 = P[24 + c];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                P[(6 + k) * 12 + c] += p0 * r9[k] + p1 * r9[3 + k] + p2 * r9[6 + k];
                P[(9 + k) * 12 + c] += dt * P[(3 + k) * 12 + c];
            }
        }
    }

    // ---- backward Riccati sweep (factors  blockdiag(Rt) + B' (state cost) B  in stage-wise form) and the solves with it ----
    BMPC_HD __forceinline__ bool factor() { return factor1(); }
    BMPC_HD __forceinline__ void solve(SV x) { solve1(x); }

    // ---- LB == 5: register-blocked sweep over VIRTUAL stages of one block (5 inputs) each; cost-to-go kept as a packed
    //      lower triangle (78).  A stage with two stance feet is two virtual stages: Z = A X + B_0 u_0, then X' = Z + B_1 u_1 ----
    // Per stage the only global traffic is P (read for P B, the congruence and the rank-NU update), B, Rt, and the factor
    // pieces the solves need:  Y = inv(L) F  (NU x 12) and L  (G = L L').  K = inv(L') Y is never formed:
    // K z = inv(L') (Y z),  K' g = Y' (inv(L) g),  F' inv(G) F = Y' Y.
    static BMPC_HD __forceinline__ constexpr int pk(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }

    // P <- A' P A on the packed triangle.  A = [I D; 0 I], D = dt [Rinv 0; 0 I]:  P21 += D' P11,  P22 += D' P12new + P21old D
    BMPC_HD void congruence_pk(SV P, SV ri) {
        double r9[9], p11[21], n21[6][6];
#pragma unroll
        for (int a = 0; a < 9; ++a) r9[a] = dt * ri[a];
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
                else n21[kp][c] = o[c] + dt * p11[pk(kp, c)];
                P[pk(6 + kp, c)] = n21[kp][c];
            }
#pragma unroll
            for (int k = 0; k <= kp; ++k) {
                double t1, t2;
                if (kp < 3) t1 = r9[kp] * n21[k][0] + r9[3 + kp] * n21[k][1] + r9[6 + kp] * n21[k][2];
                else t1 = dt * n21[k][kp];
                if (k < 3) t2 = o[0] * r9[k] + o[1] * r9[3 + k] + o[2] * r9[6 + k];
                else t2 = dt * o[k];
                P[pk(6 + kp, 6 + k)] += t1 + t2;
            }
        }
    }

    BMPC_HD __noinline__ bool factor1() {
        static_assert(LB == 5, "register-blocked sweep is written for 5 free components per block");
        SV P = ws + L::o_P;
#pragma unroll 1
        for (int e = 0; e < 78; ++e) P[e] = 0.0;
#pragma unroll
        for (int a = 0; a < 12; ++a) P[pk(a, a)] = p.Q[a];
#pragma unroll 1
        for (int v = S - 1; v >= 0; --v) {
            // virtual stage v = block (stage st, foot slot li): the first block of a stage carries the dynamics A_st and the
            // state cost of the state before it, the others act on the intermediate state (A = I, no cost)
            const int st = v / NF;
            const bool dyn = (v - st * NF) == 0;
            const double dte = dyn ? dt : 0.0;
            SV Bv = ws + (L::o_Bm + 30 * v), Rt = ws + (L::o_Rt + 25 * v), ri = ws + (L::o_rinv + 9 * st);
            SV Y = ws + (L::o_K + 60 * v), Lc = ws + (L::o_Lc + 25 * v);
            double B[6][5], lo[6][5], Lr[5][5];
#pragma unroll
            for (int k = 0; k < 6; ++k)
#pragma unroll
                for (int c = 0; c < 5; ++c) B[k][c] = Bv[k * 5 + c];
            // rows 6..11 of P B (they also feed G); rows 0..5 are produced one at a time below so that they are never all live
#pragma unroll
            for (int r = 6; r < 12; ++r) {
                double pr[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) pr[k] = P[pk(r, 6 + k)];
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) acc += pr[k] * B[k][c];
                    lo[r - 6][c] = acc;
                }
            }
            // G = Rt + B' PB[6:12], Cholesky in registers (reciprocal diagonal)
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
#pragma unroll
                for (int a = j; a < 5; ++a) {
                    double v = Rt[a * 5 + j];
#pragma unroll
                    for (int k = 0; k < 6; ++k) v += B[k][a] * lo[k][j];
#pragma unroll
                    for (int k = 0; k < j; ++k) v -= Lr[a][k] * Lr[j][k];
                    if (a == j) {
                        ok = ok && (v > 0.0) && (v < 1e300);
                        Lr[j][j] = 1.0 / sqrt(v);
                    } else {
                        Lr[a][j] = v * Lr[j][j];
                    }
                }
            }
            if (!ok) return false;
            // F = PB' A: row j < 6 of P B is column j of F and also adds into columns 6..11 (kept in lo); Y = inv(L) F
            {
                double r9[9];
#pragma unroll
                for (int a = 0; a < 9; ++a) r9[a] = dte * ri[a];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    double pr[6], h[5];
#pragma unroll
                    for (int k = 0; k < 6; ++k) pr[k] = P[pk(j, 6 + k)];
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < 6; ++k) acc += pr[k] * B[k][c];
                        h[c] = acc;
                        if (j < 3) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) lo[k][c] += acc * r9[3 * j + k];
                        } else {
                            lo[j][c] += dte * acc;
                        }
                    }
#pragma unroll
                    for (int a = 0; a < 5; ++a) {
                        double v = h[a];
#pragma unroll
                        for (int k = 0; k < a; ++k) v -= Lr[a][k] * h[k];
                        h[a] = v * Lr[a][a];
                        Y[a * 12 + j] = h[a];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 6; ++j)
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    double v = lo[j][a];
#pragma unroll
                    for (int k = 0; k < a; ++k) v -= Lr[a][k] * lo[j][k];
                    lo[j][a] = v * Lr[a][a];
                    Y[a * 12 + 6 + j] = lo[j][a];
                }
#pragma unroll
            for (int a = 0; a < 5; ++a)
#pragma unroll
                for (int k = 0; k <= a; ++k) Lc[a * 5 + k] = Lr[a][k];
            if (v == 0) break;
            // P <- [Q] + A' P A - Y' Y
            if (dyn) congruence_pk(P, ri);
            {
                double Yr[5][12];
#pragma unroll
                for (int a = 0; a < 5; ++a)
#pragma unroll
                    for (int j = 0; j < 12; ++j) Yr[a][j] = Y[a * 12 + j];
#pragma unroll
                for (int r = 0; r < 12; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) {
                        double acc = P[pk(r, c)];
#pragma unroll
                        for (int a = 0; a < 5; ++a) acc -= Yr[a][r] * Yr[a][c];
                        if (r == c && dyn) acc += p.Q[r];
                        P[pk(r, c)] = acc;
                    }
            }
        }
        return true;
    }

    BMPC_HD __noinline__ void solve1(SV x) {
        double pv[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) pv[a] = 0.0;
#pragma unroll 1
        for (int v = S - 1; v >= 0; --v) {
            const int st = v / NF;
            const double dte = ((v - st * NF) == 0) ? dt : 0.0;
            SV Bv = ws + (L::o_Bm + 30 * v), ri = ws + (L::o_rinv + 9 * st), Y = ws + (L::o_K + 60 * v), Lc = ws + (L::o_Lc + 25 * v),
               xi = x + 5 * v;
            double w[5];
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                double g = -xi[c];
#pragma unroll
                for (int k = 0; k < 6; ++k) g += Bv[k * 5 + c] * pv[6 + k];
#pragma unroll
                for (int k = 0; k < c; ++k) g -= Lc[c * 5 + k] * w[k];
                w[c] = g * Lc[c * 5 + c];
                xi[c] = w[c];
            }
            double np[12];
#pragma unroll
            for (int a = 0; a < 12; ++a) np[a] = pv[a];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                np[6 + k] += dte * (ri[k] * pv[0] + ri[3 + k] * pv[1] + ri[6 + k] * pv[2]);
                np[9 + k] += dte * pv[3 + k];
            }
#pragma unroll
            for (int a = 0; a < 5; ++a)
#pragma unroll
                for (int j = 0; j < 12; ++j) np[j] -= Y.lds(a * 12 + j) * w[a];
#pragma unroll
            for (int a = 0; a < 12; ++a) pv[a] = np[a];
        }
        double z[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) z[a] = 0.0;
#pragma unroll 1
        for (int v = 0; v < S; ++v) {
            const int st = v / NF;
            const double dte = ((v - st * NF) == 0) ? dt : 0.0;
            SV Bv = ws + (L::o_Bm + 30 * v), ri = ws + (L::o_rinv + 9 * st), Y = ws + (L::o_K + 60 * v), Lc = ws + (L::o_Lc + 25 * v),
               xi = x + 5 * v;
            double t[5], xs[5];
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                double v = xi[a];
#pragma unroll
                for (int j = 0; j < 12; ++j) v += Y.lds(a * 12 + j) * z[j];
                t[a] = v;
            }
#pragma unroll
            for (int a = 4; a >= 0; --a) {
                double v = t[a];
#pragma unroll
                for (int k = a + 1; k < 5; ++k) v -= Lc[k * 5 + a] * xs[k];
                xs[a] = v * Lc[a * 5 + a];
            }
            double acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                xs[a] = -xs[a];
                xi[a] = xs[a];
#pragma unroll
                for (int k = 0; k < 6; ++k) acc[k] += Bv[k * 5 + a] * xs[a];
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                z[a] += dte * (ri[3 * a] * z[6] + ri[3 * a + 1] * z[7] + ri[3 * a + 2] * z[8]);
                z[3 + a] += dte * z[9 + a];
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) z[6 + k] += acc[k];
        }
    }

    // -dv / v in FP32: only a step LENGTH, cut by step_frac afterwards (same as the warp-per-robot kernel); NaN/Inf propagate
    static BMPC_HD __forceinline__ float sratio(double dv, double v) { return -(float)dv / (float)v; }
    BMPC_HD __forceinline__ void crow(int k, double (&cb)[LB]) const {
#pragma unroll
        for (int c = 0; c < LB; ++c) cb[c] = Cb[k * LB + c];
    }
    BMPC_HD __forceinline__ double cdotr(int k, const double (&v)[LB]) const {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < LB; ++c) acc += Cb[k * LB + c] * v[c];
        return acc;
    }
    BMPC_HD __forceinline__ double cdot(int k, SV v) const {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < LB; ++c) acc += Cb[k * LB + c] * v[c];
        return acc;
    }
    // out_i = sign * (base_i) + (C' w)_i
    BMPC_HD void gather(SV w, SV base, double bscale, SV out) {
#pragma unroll 1
        for (int j = 0; j < S; ++j) {
            double acc[LB];
#pragma unroll
            for (int c = 0; c < LB; ++c) acc[c] = 0.0;
#pragma unroll 1
            for (int k = 0; k < mb; ++k) {
                const double wk = w[j * mb + k];
                double cb[LB];
                crow(k, cb);
#pragma unroll
                for (int c = 0; c < LB; ++c) acc[c] += cb[c] * wk;
            }
#pragma unroll
            for (int c = 0; c < LB; ++c) out[j * LB + c] = (bscale != 0.0 ? bscale * base[j * LB + c] : 0.0) + acc[c];
        }
    }

    // ---- the whole tick -------------------------------------------------------------------------------------------
    BMPC_HD void run(const IoPtrs& io, int inst) {
        mb = p.mb;
        m = S * mb;
        dt = p.dt;
        SV uv = ws + L::o_u, duv = ws + L::o_du, xv = ws + L::o_x, upv = ws + L::o_up, rdv = ws + L::o_rd, ppv = ws + L::o_pp,
           tvp = ws + L::o_tv, hd = ws + L::o_hd;
        SV r_s = ws + L::o_rs, r_l = ws + L::o_rl, r_d = ws + L::o_rdw, r_p = ws + L::o_rp, r_c = ws + L::o_rc, r_w = ws + L::o_rw;
        SV Nn = ws + L::o_Nn;

        // ---- 0. inputs; this path takes robots with exactly NF stance feet in every stage ----
        bool bad = mb > L::MBM;
#pragma unroll
        for (int a = 0; a < 12; ++a) {
            x_fb[a] = io.x_fb[(size_t)inst * 12 + a];
            bad = bad || !isfinite(x_fb[a]);
        }
        double foot[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            foot[a] = io.foot[(size_t)inst * 6 + a];
            bad = bad || !isfinite(foot[a]);
        }
        int cont0[2];
        int blockOf[2 * HZ];
        {
            int j = 0;
#pragma unroll 1
            for (int s = 0; s < HZ; ++s) {
                int cnt = 0;
                for (int l = 0; l < 2; ++l) {
                    const int c = io.contact[(size_t)inst * 2 * HZ + 2 * s + l] ? 1 : 0;
                    if (s == 0) cont0[l] = c;
                    blockOf[2 * s + l] = -1;
                    if (c) {
                        if (cnt < NF) fo[j] = l, blockOf[2 * s + l] = j, ++j;
                        ++cnt;
                    }
                }
                if (cnt != NF) bad = true, j = (s + 1) * NF;
            }
        }
        if (bad) {  // not this path's robot (or bad input): the warp-per-robot kernels take it
            io.status[inst] = 1;
            io.iters[inst] = 0;
            return;
        }
        const int phase_k = io.phase_k[inst];

        // ---- 1. references, per-stage dynamics and input maps (MPC.py:61-109, 148-185) ----
        double rotn[9], footv[18], ub[LB];
        {
            const double hh = (double)p.h;
            const double ex = p.kv * (x_fb[3] - p.x_cmd[3]), ey = p.kv * (x_fb[4] - p.x_cmd[4]);
            const double x1 = x_fb[3] + x_fb[9] * 1 / 2 * hh / 2 * dt + ex;
            const double x2 = x_fb[3] + x_fb[9] * 1 / 2 * hh * dt + ex;
            const double y1 = x_fb[4] + x_fb[10] * 1 / 2 * hh / 2 * dt + ey;
            const double y2 = x_fb[10] + x_fb[10] * 1 / 2 * hh * dt + ey;  // MPC.py:87 starts from x_fb[10]
            for (int c = 0; c < 6; ++c) footv[c] = foot[c];
            footv[6] = x1, footv[7] = y1, footv[8] = 0.0, footv[9] = x1, footv[10] = y1, footv[11] = 0.0;
            footv[12] = x2, footv[13] = y2, footv[14] = 0.0, footv[15] = x2, footv[16] = y2, footv[17] = 0.0;
            eul2rotm(x_fb, rotn);
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
            for (int l = 0; l < 2; ++l)
                for (int c = 0; c < LB; ++c) {
                    const int ca = p.comps[c];
                    Rd[l][c] = p.R[(ca < 3) ? (3 * l + ca) : (6 + 3 * l + ca - 3)];
                }
        }
        const double vm = dt / p.mass;
        bool singular = false;
#pragma unroll 1
        for (int k = 0; k < HZ; ++k) {
            const int kk = phase_k % 5;
            int sel = 0;
            if (cont0[0] + cont0[1] == 1) sel = (k < 5 - kk) ? 0 : ((k < 10 - kk) ? 1 : 2);
            double xr[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) xr[i] = (k == 0) ? x_fb[i] : p.x_cmd[i];
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    if (p.x_cmd[i + 6] != 0.0) xr[i] = x_fb[i] + p.x_cmd[i + 6] * (k * dt);
            }
            SV xref = ws + (L::o_xref + 12 * k);
#pragma unroll
            for (int i = 0; i < 12; ++i) xref[i] = xr[i];
            double sz, cz, sy, cy, sx, cx;  // dynamics read x[0] as yaw, x[1] pitch, x[2] roll (MPC.py:151-153)
            sincos(xr[0], &sz, &cz);
            sincos(xr[1], &sy, &cy);
            sincos(xr[2], &sx, &cx);
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
            SV ri = ws + (L::o_rinv + 9 * k);
            ri[0] = cz * icp, ri[1] = sz * icp, ri[2] = 0.0;
            ri[3] = -sz, ri[4] = cz, ri[5] = 0.0;
            ri[6] = cz * sy * icp, ri[7] = sz * sy * icp, ri[8] = 1.0;
            double cw[6] = {0, 0, 0, 0, 0, 0};
            SV B0 = ws + (L::o_B0 + 6 * NU * k);
            for (int li = 0; li < NF; ++li) {
                const int l = fo[k * NF + li];
                const double* fr = footv + 6 * sel + 3 * l;
                const double r0 = fr[0] - xr[3], r1 = fr[1] - xr[4], r2 = fr[2] - xr[5];
                double B[18];  // dt Iw^{-1} [skew(r) | I]  (MPC.py:174-179, 184)
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const double i0 = ii[3 * a], i1 = ii[3 * a + 1], i2 = ii[3 * a + 2];
                    B[6 * a + 0] = dt * (i1 * r2 - i2 * r1);
                    B[6 * a + 1] = dt * (i2 * r0 - i0 * r2);
                    B[6 * a + 2] = dt * (i0 * r1 - i1 * r0);
                    B[6 * a + 3] = dt * i0, B[6 * a + 4] = dt * i1, B[6 * a + 5] = dt * i2;
                }
                for (int a = 0; a < 3; ++a) {
                    for (int c = 0; c < LB; ++c) {
                        B0[a * NU + li * LB + c] = B[6 * a + p.comps[c]];
                        B0[(3 + a) * NU + li * LB + c] = (p.comps[c] == a) ? vm : 0.0;
                    }
                    for (int c = 0; c < p.npinned; ++c) {
                        cw[a] += B[6 * a + p.pinned[c]] * p.lo6[p.pinned[c]];
                        if (p.pinned[c] == a) cw[3 + a] += vm * p.lo6[p.pinned[c]];
                    }
                }
            }
            cw[5] -= p.g * dt;
            SV cc = ws + (L::o_c + 6 * k);
#pragma unroll
            for (int a = 0; a < 6; ++a) cc[a] = cw[a];
        }
        if (singular) {
            io.status[inst] = 1;
            io.iters[inst] = 0;
            return;
        }
        // per-block inequality rows in block coordinates (MPC.py:220-271), the same for every block
        for (int k = 0; k < mb; ++k) {
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
                f6[0] = -len * rotn[2], f6[1] = -len * rotn[5], f6[2] = -len * rotn[8];
                f6[3] = sg * rotn[1], f6[4] = sg * rotn[4], f6[5] = sg * rotn[7];
            }
            for (int c = 0; c < p.npinned; ++c) rhs -= f6[p.pinned[c]] * p.lo6[p.pinned[c]];
            for (int c = 0; c < LB; ++c) Cb[k * LB + c] = f6[p.comps[c]];
            rb[k] = rhs;
        }

        // ---- 2. interior point ----
#pragma unroll 1
        for (int i = 0; i < N; ++i) uv[i] = 0.0;
        grad(uv, tvp, nullptr);  // g = gradient at u = 0
        double gs = 0.0;
#pragma unroll 1
        for (int i = 0; i < N; ++i) gs = fmax(gs, fabs(tvp[i]));
        gs += 1.0;
#pragma unroll 1
        for (int i = 0; i < N; ++i) uv[i] = ub[i % LB];
        double part = 0.0;
#pragma unroll 1
        for (int k = 0; k < mb; ++k) {
            double sl = rb[k];
            for (int c = 0; c < LB; ++c) sl -= Cb[k * LB + c] * ub[c];
            if (!(sl > 1e-3)) sl = 1.0;
            for (int j = 0; j < S; ++j) r_s[j * mb + k] = sl;
            part += sl * (double)S;
        }
        const double mu0 = p.mu0_scale * part / (double)m;
#pragma unroll 1
        for (int r = 0; r < m; ++r) r_l.sts(r, mu0 / r_s.lds(r));
#pragma unroll 1
        for (int j = 0; j < S; ++j) {  // input maps in use, block layout [block][6][LB] (the problem's maps B0 are [stage][6][NU])
            const int st = j / NF, li = j - st * NF;
            for (int k = 0; k < 6; ++k)
                for (int c = 0; c < LB; ++c) ws[L::o_Bm + (j * 6 + k) * LB + c] = ws[L::o_B0 + 6 * NU * st + k * NU + li * LB + c];
        }

        const double mu_target = p.mu_tol * gs;
        int status = 1, it = 0;
        double mu = 0.0, rdmax = 0.0;
        bool rd_fresh = false;
        double alpha_prev = 0.0;
#pragma unroll 1
        while (true) {
            if (it >= p.max_iter) {
                status = 1;
                break;
            }
            ++it;
            if (!rd_fresh) {
                grad(uv, tvp, nullptr);
                gather(r_l, tvp, 1.0, rdv);
                rdmax = 0.0;
#pragma unroll 1
                for (int i = 0; i < N; ++i) rdmax = fmax(rdmax, fabs(rdv[i]));
                rd_fresh = true;
            }
            // One pass per block: barrier weights d = lam / s, primal residuals, the predictor right-hand side
            // -rd - C'(d rp - lam) and the stage input weights R + Cb' diag(d) Cb (written where the sweep reads them)
            part = 0.0;
#pragma unroll 1
            for (int j = 0; j < S; ++j) {
                double vb[LB], gacc[LB], acc[LB * (LB + 1) / 2];
#pragma unroll
                for (int c = 0; c < LB; ++c) {  // the previous iteration's step is applied here (alpha_prev = 0 in the first one)
                    vb[c] = uv[j * LB + c];
                    gacc[c] = rdv[j * LB + c];
                    if (alpha_prev != 0.0) {  // (duv is uninitialised before the first step)
                        vb[c] += alpha_prev * duv[j * LB + c];
                        gacc[c] *= (1.0 - alpha_prev);
                        uv[j * LB + c] = vb[c];
                        rdv[j * LB + c] = gacc[c];
                    }
                }
#pragma unroll
                for (int a = 0; a < LB; ++a)
#pragma unroll
                    for (int b2 = 0; b2 <= a; ++b2) acc[a * (a + 1) / 2 + b2] = (a == b2) ? Rd[fo[j]][a] : 0.0;
#pragma unroll 1
                for (int k = 0; k < mb; ++k) {
                    const int r = j * mb + k;
                    double s = r_s.lds(r), l = r_l.lds(r);
                    if (alpha_prev != 0.0) {
                        s += alpha_prev * r_p.lds(r);
                        l += alpha_prev * r_c.lds(r);
                        r_s.sts(r, s);
                        r_l.sts(r, l);
                    }
                    double cb[LB];
                    crow(k, cb);
                    double cu = 0.0;
#pragma unroll
                    for (int c = 0; c < LB; ++c) cu += cb[c] * vb[c];
                    const double d = l / s, rp = cu + s - rb[k];
                    r_p.sts(r, rp);
                    const double w = d * rp - l;
                    part += s * l;
#pragma unroll
                    for (int c = 0; c < LB; ++c) gacc[c] += cb[c] * w;
#pragma unroll
                    for (int a = 0; a < LB; ++a)
#pragma unroll
                        for (int b2 = 0; b2 <= a; ++b2) acc[a * (a + 1) / 2 + b2] += cb[a] * cb[b2] * d;
                }
#pragma unroll
                for (int c = 0; c < LB; ++c) xv[j * LB + c] = -gacc[c];
                SV Rt = ws + (L::o_Rt + LB * LB * j);  // block layout: the sweep reads one LB x LB block per virtual stage
#pragma unroll
                for (int a = 0; a < LB; ++a)
#pragma unroll
                    for (int b2 = 0; b2 <= a; ++b2) Rt[a * LB + b2] = acc[a * (a + 1) / 2 + b2];
            }
            mu = part / (double)m;
            if (mu <= mu_target && rdmax <= p.rd_tol * mu_target) {
                status = 0;
                break;
            }
            if (!factor()) {
                status = 2;
                break;
            }
            solve(xv);
            float ratio = 0.f;
            part = 0.0;
#pragma unroll 1
            for (int j = 0; j < S; ++j) {
                double vb[LB];
#pragma unroll
                for (int c = 0; c < LB; ++c) vb[c] = xv[j * LB + c];
#pragma unroll 1
                for (int k = 0; k < mb; ++k) {
                    const int r = j * mb + k;
                    const double sr = r_s.lds(r), lr = r_l.lds(r);
                    const double dsa = -r_p.lds(r) - cdotr(k, vb);
                    const double dla = -lr - (lr / sr) * dsa;
                    ratio = fmaxf(ratio, fmaxf(sratio(dsa, sr), sratio(dla, lr)));
                    r_c.sts(r, dsa * dla);
                    part += dsa * dla;
                }
            }
            const double a_aff = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
            const double mu_aff = mu * (1.0 - a_aff) + a_aff * a_aff * part / (double)m;
            double sigma = mu_aff / mu;
            sigma = sigma * sigma * sigma;
            const double tgt = sigma * mu;
            // corrector right-hand side C' wc, wc = (dsa dla - sigma mu) / s   (solved in duv, the affine step stays in xv)
#pragma unroll 1
            for (int j = 0; j < S; ++j) {
                double gacc[LB];
#pragma unroll
                for (int c = 0; c < LB; ++c) gacc[c] = 0.0;
#pragma unroll 1
                for (int k = 0; k < mb; ++k) {
                    const int r = j * mb + k;
                    const double wc = (r_c.lds(r) - tgt) / r_s.lds(r);
                    r_c.sts(r, wc);
                    double cb[LB];
                    crow(k, cb);
#pragma unroll
                    for (int c = 0; c < LB; ++c) gacc[c] += cb[c] * wc;
                }
#pragma unroll
                for (int c = 0; c < LB; ++c) duv[j * LB + c] = gacc[c];
            }
            solve(duv);
#pragma unroll 1
            for (int i = 0; i < N; ++i) duv[i] += xv[i];
            ratio = 0.f;
#pragma unroll 1
            for (int j = 0; j < S; ++j) {
                double vb[LB];
#pragma unroll
                for (int c = 0; c < LB; ++c) vb[c] = duv[j * LB + c];
#pragma unroll 1
                for (int k = 0; k < mb; ++k) {
                    const int r = j * mb + k;
                    const double sr = r_s.lds(r), lr = r_l.lds(r);
                    const double ds = -r_p.lds(r) - cdotr(k, vb);
                    const double dl = -lr - r_c.lds(r) - (lr / sr) * ds;
                    ratio = fmaxf(ratio, fmaxf(sratio(ds, sr), sratio(dl, lr)));
                    r_p.sts(r, ds);
                    r_c.sts(r, dl);
                }
            }
            double a2 = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
            if (!isfinite(ratio)) {
                status = 2;
                break;
            }
            if (kGondzio && p.gondzio && a2 < p.gondzio_below) {
                const double at = fmin(1.0, 1.5 * a2 + 0.1);
#pragma unroll 1
                for (int r = 0; r < m; ++r) {
                    const double v = (r_s.lds(r) + at * r_p.lds(r)) * (r_l.lds(r) + at * r_c.lds(r));
                    double vt = fmin(fmax(v, 0.1 * tgt), 10.0 * tgt) - v;
                    vt = fmax(vt, -10.0 * tgt);
                    r_w[r] = -vt / r_s.lds(r);
                }
                gather(r_w, xv, 0.0, xv);
                solve(xv);
                ratio = 0.f;
#pragma unroll 1
                for (int j = 0; j < S; ++j) {
                    double vb[LB];
#pragma unroll
                    for (int c = 0; c < LB; ++c) vb[c] = xv[j * LB + c];
#pragma unroll 1
                    for (int k = 0; k < mb; ++k) {
                        const int r = j * mb + k;
                        const double cx = cdotr(k, vb);
                        const double ds = r_p.lds(r) - cx;
                        const double dl = r_c.lds(r) + r_d[r] * cx - r_w[r];
                        ratio = fmaxf(ratio, fmaxf(sratio(ds, r_s.lds(r)), sratio(dl, r_l.lds(r))));
                    }
                }
                const double a3 = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
                if (isfinite(ratio) && a3 > a2) {
                    a2 = a3;
#pragma unroll 1
                    for (int j = 0; j < S; ++j) {
                        double vb[LB];
#pragma unroll
                        for (int c = 0; c < LB; ++c) vb[c] = xv[j * LB + c];
#pragma unroll 1
                        for (int k = 0; k < mb; ++k) {
                            const int r = j * mb + k;
                            const double cx = cdotr(k, vb);
                            r_p.sts(r, r_p.lds(r) - (cx));
                            r_c.sts(r, r_c.lds(r) + (r_d[r] * cx - r_w[r]));
                        }
                    }
#pragma unroll 1
                    for (int i = 0; i < N; ++i) duv[i] += xv[i];
                }
            }
            // the step (u, s, lam, rd) is applied by the next iteration's first pass
            alpha_prev = (it > 14 ? 0.9 : p.step_frac) * a2;
            rdmax *= (1.0 - alpha_prev);
        }

        // ---- 3. active-set polish + certificate (one attempt; anything else goes to the warp-per-robot kernels) ----
        int amask[S], bdim[S];
        bool polished = false;
        if (status == 0) {
            // diag(Hc) from the uncontrolled cost-to-go
            {
                SV P = ws + L::o_P;
#pragma unroll 1
                for (int e = 0; e < 144; ++e) P[e] = 0.0;
#pragma unroll
                for (int a = 0; a < 12; ++a) P[a * 13] = p.Q[a];
#pragma unroll 1
                for (int i = HZ - 1; i >= 0; --i) {
                    SV B = ws + (L::o_B0 + 6 * NU * i);
#pragma unroll 1
                    for (int c = 0; c < NU; ++c) {
                        double b[6], acc = Rd[fo[i * NF + c / LB]][c % LB];
#pragma unroll
                        for (int k = 0; k < 6; ++k) b[k] = B[k * NU + c];
#pragma unroll
                        for (int k = 0; k < 6; ++k)
#pragma unroll
                            for (int k2 = 0; k2 < 6; ++k2) acc += b[k] * P[(6 + k) * 12 + 6 + k2] * b[k2];
                        hd[i * NU + c] = acc;
                    }
                    if (i == 0) break;
                    congruence(P, ws + (L::o_rinv + 9 * i));
#pragma unroll
                    for (int a = 0; a < 12; ++a) P[a * 13] += p.Q[a];
                }
            }
#pragma unroll 1
            for (int j = 0; j < S; ++j) {
                int mk = 0;
#pragma unroll 1
                for (int k = 0; k < mb; ++k) {
                    double th = 0.0, aa = 0.0;
#pragma unroll
                    for (int c = 0; c < LB; ++c) th += Cb[k * LB + c] * Cb[k * LB + c] * hd[j * LB + c], aa += Cb[k * LB + c] * Cb[k * LB + c];
                    if (r_l[j * mb + k] * fmax(aa * aa, 1e-300) > th * r_s[j * mb + k]) mk |= 1 << k;
                }
                amask[j] = mk;
            }
            status = 1;
#pragma unroll 1
            for (int round = 0; round < p.polish_rounds; ++round) {
                bool bad_blk = false;
#pragma unroll 1
                for (int j = 0; j < S; ++j) {
                    double pl[LB], Nl[E];
                    int dim = 0;
                    if (!block_nullspace<LB>(Cb, rb, mb, (unsigned)amask[j], pl, Nl, &dim)) bad_blk = true;
                    bdim[j] = dim;
                    for (int c = 0; c < LB; ++c) ppv[j * LB + c] = pl[c];
                    for (int e = 0; e < E; ++e) Nn[j * E + e] = Nl[e];
                }
                if (bad_blk) break;
                grad(ppv, tvp, nullptr);
                // reduced LQR: inputs w_b, maps B_b N_b, weights N_b' R N_b (+ I on the padding)
#pragma unroll 1
                for (int s = 0; s < HZ; ++s) {
                    SV B0 = ws + (L::o_B0 + 6 * NU * s);
                    for (int li = 0; li < NF; ++li) {
                        const int j = s * NF + li, dim = bdim[j];
                        SV Nj = Nn + j * E, Rt = ws + (L::o_Rt + LB * LB * j), Bm = ws + (L::o_Bm + 6 * LB * j);  // block layouts
#pragma unroll 1
                        for (int a = 0; a < LB; ++a) {
                            double acc = 0.0;
#pragma unroll
                            for (int c = 0; c < LB; ++c) acc += Nj[c * LB + a] * tvp[j * LB + c];
                            xv[j * LB + a] = (a < dim) ? -acc : 0.0;
#pragma unroll 1
                            for (int b = 0; b <= a; ++b) {
                                double w = 0.0;
#pragma unroll
                                for (int c = 0; c < LB; ++c) w += Nj[c * LB + a] * Rd[fo[j]][c] * Nj[c * LB + b];
                                if (a == b && a >= dim) w = 1.0;
                                Rt[a * LB + b] = w;
                            }
#pragma unroll
                            for (int k = 0; k < 6; ++k) {
                                double w = 0.0;
#pragma unroll
                                for (int c = 0; c < LB; ++c) w += B0[k * NU + li * LB + c] * Nj[c * LB + a];
                                Bm[k * LB + a] = w;
                            }
                        }
                    }
                }
                if (!factor()) break;
                solve(xv);
#pragma unroll 1
                for (int j = 0; j < S; ++j) {
                    SV Nj = Nn + j * E;
#pragma unroll 1
                    for (int c = 0; c < LB; ++c) {
                        double acc = ppv[j * LB + c];
#pragma unroll
                        for (int a = 0; a < LB; ++a) acc += Nj[c * LB + a] * xv[j * LB + a];
                        upv[j * LB + c] = acc;
                    }
                }
                // primal check: violated inactive rows join the active set
                bool changed = false;
#pragma unroll 1
                for (int j = 0; j < S; ++j) {
                    double vb[LB];
#pragma unroll
                    for (int c = 0; c < LB; ++c) vb[c] = upv[j * LB + c];
#pragma unroll 1
                    for (int k = 0; k < mb; ++k) {
                        const double bk = rb[k];
                        const double viol = cdotr(k, vb) - bk;
                        if (viol > 1e-9 * (1.0 + fabs(bk)) && !((amask[j] >> k) & 1)) amask[j] |= 1 << k, changed = true;
                    }
                }
                if (changed) continue;
                // dual check: minus the gradient must be a non-negative combination of the active rows
                grad(upv, tvp, nullptr);
                bool fail = false;
#pragma unroll 1
                for (int j = 0; j < S; ++j) {
                    double rneg[LB], lam[L::MBM];
#pragma unroll
                    for (int c = 0; c < LB; ++c) rneg[c] = -tvp[j * LB + c];
                    for (int k = 0; k < mb; ++k) lam[k] = r_l[j * mb + k];
                    unsigned drop = 0u;
                    if (block_dual_fast<LB>(Cb, mb, (unsigned)amask[j], lam, rneg, gs)) continue;
                    if (!block_dual_check<LB>(Cb, mb, (unsigned)amask[j], rneg, gs, &drop)) {
                        if (drop == 0u) fail = true;
                        else amask[j] &= ~(int)drop, changed = true;
                    }
                }
                if (fail) break;
                if (!changed) {
                    polished = true;
                    break;
                }
            }
        }
        if (!polished) {
            io.status[inst] = 1;
            io.iters[inst] = it;
            return;
        }

        // ---- 4. outputs ----
        double u0[12];
        double umax = 0.0;
        for (int e = 0; e < HZ * 12; ++e) {
            const int s = e / 12, c12 = e - 12 * s;
            const int l = (c12 % 6) / 3, comp = (c12 < 6) ? (c12 % 3) : (3 + c12 % 3);
            const int b = blockOf[2 * s + l];
            double val = 0.0;
            if (b >= 0) {
                val = p.lo6[comp];
                for (int c = 0; c < LB; ++c)
                    if (p.comps[c] == comp) val = upv[b * LB + c];
            }
            io.controls[(size_t)inst * HZ * 12 + e] = val;
            umax = fmax(umax, fabs(val));
            if (s == 0) u0[c12] = val;
        }
        const double uscale = fmax(1.0, umax);
        if (io.states) grad(upv, tvp, io.states + (size_t)inst * HZ * 13);
        if (io.fric_active) {
            for (int s = 0; s < HZ; ++s) {
                const double tol = 1e-6 * uscale;
                unsigned mask = 0;
                for (int l = 0; l < 2; ++l) {
                    const int b = blockOf[2 * s + l];
                    if (b < 0) continue;
                    double f[3] = {p.lo6[0], p.lo6[1], p.lo6[2]};
                    for (int c = 0; c < LB; ++c)
                        if (p.comps[c] < 3) f[p.comps[c]] = upv[b * LB + c];
                    if (f[2] <= tol) continue;
                    const double res[4] = {f[0] - p.mu * f[2], f[1] - p.mu * f[2], -f[0] - p.mu * f[2], -f[1] - p.mu * f[2]};
                    for (int r = 0; r < 4; ++r)
                        if (res[r] >= -tol) mask |= 1u << (4 * l + r);
                }
                io.fric_active[(size_t)inst * HZ + s] = (uint8_t)mask;
            }
        }
        if (io.do_lowlevel && io.tau) {
            double q[10], qd[10], pf[6];
            for (int a = 0; a < 10; ++a) q[a] = io.q[(size_t)inst * 10 + a], qd[a] = io.qd[(size_t)inst * 10 + a];
            for (int a = 0; a < 6; ++a) pf[a] = io.pf_w[(size_t)inst * 6 + a];
            for (int leg = 0; leg < 2; ++leg) {
                double tl[5];
                lowlevel_leg(p, x_fb, io.t_swing[inst], pf, q, qd, rotn, leg, (double)cont0[leg], u0, tl);
                for (int c = 0; c < 5; ++c) io.tau[(size_t)inst * 10 + 5 * leg + c] = tl[c];
            }
        }
        if (io.ws_mask)
            for (int e = 0; e < 2 * HZ; ++e) io.ws_mask[(size_t)inst * 2 * HZ + e] = blockOf[e] >= 0 ? amask[blockOf[e]] : -1;
        io.status[inst] = 0;
        io.iters[inst] = it;
        if (io.resid) io.resid[2 * inst] = 0.0, io.resid[2 * inst + 1] = rdmax;
    }
};

#if defined(__CUDACC__) && !defined(BMPC_LANE_HOST_ONLY)
// One warp = 32 robots of the class's work list at a time (dynamic: a global counter hands out 32-robot slices).
#ifndef BMPC_LANE_MINB
#define BMPC_LANE_MINB 3
#endif
template <int HZ, int NF, int LB>
__global__ void __launch_bounds__(128, BMPC_LANE_MINB) lane_tick_kernel(const __grid_constant__ DevParams p, const IoPtrs io,
                                                        const int* __restrict__ work_list, const int* __restrict__ work_count,
                                                        int* __restrict__ slice_counter, double* __restrict__ wsbase, int min_count) {
    using L = LaneL<HZ, NF, LB>;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int count = *work_count;
    // a class too small to fill the machine with 32-robot slices is faster on the warp-per-robot kernel (one slice takes
    // ~20 ms whatever the batch): leave it alone, collect_or_all_kernel then hands the whole list over
    if (count < min_count) return;
    SV ws{wsbase + (size_t)warp * L::total * 32 + lane};
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(slice_counter, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        if (base + lane < count) {
            LaneSolver<HZ, NF, LB> solver(p, ws);
            solver.run(io, work_list[base + lane]);
        }
        __syncwarp();
    }
}
#endif

}  // namespace bmpc
