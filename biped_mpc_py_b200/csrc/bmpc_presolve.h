// Host-side presolve shared by the library (bmpc.cu) and the host-compiled unit test of the lane solver
// (tests/lane_host.cu): bmpc_params -> DevParams.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/biped_mpc_b200.h"
#include "bmpc_kernels.cuh"

namespace bmpc {

// Host-side presolve of the per-block inequality rows (MPC.py:220-271): which of the six
// components [fx,fy,fz,mx,my,mz] are free, and which rows are implied by others for these
// parameter values (dropping them does not change the feasible set).
inline int build_dev_params_impl(const bmpc_params& P, DevParams& d, std::string& err) {
    auto fail = [&](const char* m) { err = m; return 1; };
    memset(&d, 0, sizeof(d));
    if (P.h != 10 && P.h != 30) return fail("horizon must be 10 or 30 (the instantiated kernels)");
    d.h = P.h;
    d.extend = P.extend_gait;
    d.dt = P.dt;
    d.kv = P.kv;
    d.swing_height = P.swing_height;
    d.mass = P.mass;
    d.lt_eff = P.lt - 0.01;  // MPC.py:254
    d.lh_eff = P.lh - 0.02;  // MPC.py:255
    d.g = P.g;
    d.mu = P.mu;
    d.max_iter = P.max_iter > 0 ? P.max_iter : 40;
    d.mu_tol = P.mu_tol > 0 ? P.mu_tol : 1e-7;
    d.rd_tol = P.rd_tol > 0 ? P.rd_tol : 10.0;
    d.gondzio = 1;
    d.gondzio_below = 0.95;  // centrality corrector only when the step length is below this: 1.722 M solves/s vs 1.696 M always (131,072 robots)
    d.init_fz_frac = 0.2;   // start point: 20 % of the fz range, friction/moment components centred
    d.mu0_scale = 0.1;      // initial complementarity = mu0_scale * mean slack
    // polish rounds per attempt: long horizons have more weakly active rows that only show up as violations
    // one round at a time (h = 30 instances needing 5-7 rounds were measured with tools/kernel_model.py)
    d.polish_rounds = P.h > 10 ? 16 : 4;
    d.lock_mode = 3;        // loose lockstep of the robots of a CTA (bmpc_tick.cuh)
    // bulk L2 prefetch of the next block record: off.  It paid while the lane kernels waited on latency; their first pass now moves
    // 5.5 TB/s of its own workspace (85 % of the measured copy bandwidth) and the prefetch's unused elements only add to that:
    // 28.2 ms per 262,144-robot tick without, 28.5 - 28.9 with (bmpc_set_option "lane_prefetch" turns it back on)
    d.lane_prefetch = 0;
    d.lane_sync = 2;
    d.warm_rounds = 6;      // polish rounds allowed to a warm-started tick before it falls back to the cold path
    d.step_frac = 0.99;     // fraction of the step to the boundary (0.9 once an instance is past 14 iterations)
    memcpy(d.x_cmd, P.x_cmd, sizeof(d.x_cmd));
    memcpy(d.Q, P.Q, sizeof(d.Q));
    memcpy(d.R, P.R, sizeof(d.R));
    memcpy(d.kp, P.kp, sizeof(d.kp));
    memcpy(d.kd, P.kd, sizeof(d.kd));
    memcpy(d.inertia, P.inertia, sizeof(d.inertia));
    memcpy(d.hip, P.hip_offset, sizeof(d.hip));
    for (int c = 0; c < 3; ++c) {
        d.lo6[c] = P.f_min[c];
        d.hi6[c] = P.f_max[c];
        d.lo6[3 + c] = P.tau_min[c];
        d.hi6[3 + c] = P.tau_max[c];
    }
    if (!(P.dt > 0) || !(P.mass > 0) || !(P.mu >= 0)) return fail("dt, mass must be positive and mu non-negative");
    for (int c = 0; c < 6; ++c) {
        if (!(d.hi6[c] >= d.lo6[c])) return fail("empty box: a *_max is below its *_min");
        if (d.hi6[c] > d.lo6[c])
            d.comps[d.LB++] = c;
        else
            d.pinned[d.npinned++] = c;
    }
    if (d.LB != 5 && d.LB != 6)
        return fail("unsupported limits: at most one of the six force/moment components may be pinned (min == max)");
    bool f_free = true;
    for (int c = 0; c < 3; ++c) f_free = f_free && (d.hi6[c] > d.lo6[c]);
    auto local = [&](int comp) {
        for (int c = 0; c < d.LB; ++c)
            if (d.comps[c] == comp) return c;
        return -1;
    };
    int mb = 0;
    auto add = [&](int kind, int arg) {
        d.row_kind[mb] = kind;
        d.row_arg[mb] = arg;
        ++mb;
    };
    const bool pyramid = f_free && P.mu > 0;
    for (int c = 0; c < 6; ++c) {  // lower bounds
        if (local(c) < 0) continue;
        bool keep = true;
        if (pyramid && c == 2 && d.lo6[2] <= 0) keep = false;                       // |fx| <= mu fz => fz >= 0
        if (pyramid && c < 2 && d.lo6[c] <= -P.mu * d.hi6[2]) keep = false;          // fx >= -mu fz >= -mu fz_max
        if (keep) add(ROW_LO, local(c));
    }
    for (int c = 0; c < 6; ++c) {  // upper bounds
        if (local(c) < 0) continue;
        bool keep = true;
        if (pyramid && c < 2 && d.hi6[c] >= P.mu * d.hi6[2]) keep = false;           // fx <= mu fz <= mu fz_max
        if (keep) add(ROW_HI, local(c));
    }
    for (int r = 0; r < 4; ++r) {  // friction pyramid, MPC.py:220-229
        bool keep = true;
        // -fx - mu fz <= 0 is implied by fx >= 0 and fz >= 0 (reference defaults: f_min = 0)
        if (r >= 2 && d.lo6[r - 2] >= 0 && d.lo6[2] >= 0) keep = false;
        if (keep) add(ROW_FRIC, r);
    }
    add(ROW_LINE, 0);  // MPC.py:258-263
    add(ROW_LINE, 1);
    d.mb = mb;
    return 0;
}

}  // namespace bmpc
