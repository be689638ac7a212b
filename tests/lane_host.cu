// Host build of the lane-per-robot solver (biped_mpc_py_b200/csrc/bmpc_lane.cuh) for CPU unit tests: the SAME source
// the GPU kernel runs, one robot per call, checked against the oracle in tests/test_lane_host.py.
// Test infrastructure only: nothing in the product loads this library.
#include <vector>

#include "../biped_mpc_py_b200/csrc/bmpc_presolve.h"
#include "../biped_mpc_py_b200/csrc/bmpc_lane.cuh"
#include "../biped_mpc_py_b200/csrc/bmpc_lane_api.h"

using namespace bmpc;

// two_pass: the first call parks the stragglers (LaneDefer, bmpc_lane_api.h: a robot whose interior point has not converged
// after g_ipm_inline iterations, a robot that needs more than one polish round) and later calls continue from the parked
// records, as the launches of the GPU dispatch do
static int g_two_pass = 0, g_parked = 0, g_parked_ipm = 0, g_ipm_inline = 9;
extern "C" void lane_host_set_two_pass(int on, int ipm_inline) { g_two_pass = on, g_ipm_inline = ipm_inline; }
extern "C" int lane_host_parked(void) { return g_parked; }
extern "C" int lane_host_parked_ipm(void) { return g_parked_ipm; }

template <int HZ, int NF, bool F32, unsigned RM>
static void run_f(const DevParams& d, std::vector<double>& w, std::vector<double>& ps, const IoPtrs& io, int i) {
    if (!g_two_pass) {
        const LaneDefer none{};
        LaneSolver<HZ, NF, F32, RM>(d, SV{w.data()}, SV{ps.data()}, 0, none).run(io, i);
        return;
    }
    // one parked record per store; on the host the lane interleave is 1, but the record layout keeps its 32-wide stride
    std::vector<float> buf((size_t)LaneRec<HZ, NF>::defer_floats * 32, 0.f), ibuf((size_t)LaneRec<HZ, NF>::ipm_floats * 32, 0.f);
    int list[32] = {0}, count = 0, ilist[32] = {0}, icount = 0;
    const LaneDefer df{buf.data(), list, &count, 32, 1, 0, ibuf.data(), ilist, &icount, 32, g_ipm_inline};
    LaneSolver<HZ, NF, F32, RM> first(d, SV{w.data()}, SV{ps.data()}, 0, df);
    first.park = true;
    first.run(io, i);
    if (icount) {
        g_parked_ipm += 1;
        for (auto& x : w) x = -7.0;  // a later pass must not depend on anything an earlier one left in the workspace
        LaneSolver<HZ, NF, F32, RM> second(d, SV{w.data()}, SV{ps.data()}, 0, df);
        second.mode = 2;
        second.park = true;
        second.run(io, ilist[0], 0);
    }
    if (count) {
        g_parked += 1;
        for (auto& x : w) x = -7.0;
        LaneSolver<HZ, NF, F32, RM> third(d, SV{w.data()}, SV{ps.data()}, 0, df);
        third.mode = 1;
        third.run(io, list[0], 0);
    }
}
template <int HZ, int NF, unsigned RM>
static void run_rm(const DevParams& d, std::vector<double>& w, std::vector<double>& ps, const IoPtrs& io, int i, bool f32) {
#ifdef LANE_HOST_MINIMAL  // (sanitizer build: only what the GPU runs - float factor storage)
    (void)f32;
    run_f<HZ, NF, true, RM>(d, w, ps, io, i);
#else
    if (f32) run_f<HZ, NF, true, RM>(d, w, ps, io, i);
    else run_f<HZ, NF, false, RM>(d, w, ps, io, i);
#endif
}
// same choice as bmpc_lane.cu: the instantiation specialised for the presolve's row set if there is one, else the generic one
template <int HZ, int NF>
static void run_one(const DevParams& d, std::vector<double>& w, std::vector<double>& ps, const IoPtrs& io, int i, bool f32) {
    const unsigned rm = lane_rowmask(d);
#ifdef LANE_HOST_MINIMAL  // (sanitizer build: the reference's row set through its specialised instantiation, everything else generic)
    if (rm == kRowsRef) run_rm<HZ, NF, kRowsRef>(d, w, ps, io, i, f32);
    else run_rm<HZ, NF, 0u>(d, w, ps, io, i, f32);
#else
    if (rm == kRowsRef) run_rm<HZ, NF, kRowsRef>(d, w, ps, io, i, f32);
    else if (rm == kRowsSym) run_rm<HZ, NF, kRowsSym>(d, w, ps, io, i, f32);
    else run_rm<HZ, NF, 0u>(d, w, ps, io, i, f32);
#endif
}

// factor storage of the interior point: 1 = float (what the GPU kernels run), 0 = double (numerical reference for the tests)
static int g_f32 = 1;
extern "C" void lane_host_set_f32(int on) { g_f32 = on; }

extern "C" int lane_host_tick(const bmpc_params* P, int n, const double* x_fb, const int32_t* phase_k, const double* t_swing,
                              const double* foot, const uint8_t* contact, const double* q, const double* qd, const double* pf_w,
                              double* controls, double* states, double* tau, int32_t* status, int32_t* iters,
                              uint8_t* fric_active, double* resid, int32_t* ws_mask) {
    DevParams d;
    std::string err;
    if (build_dev_params_impl(*P, d, err)) return 1;
    if ((d.h != 10 && d.h != 30) || d.LB != 5) return 2;
    IoPtrs io{};
    io.x_fb = x_fb, io.phase_k = phase_k, io.t_swing = t_swing, io.foot = foot, io.contact = contact, io.q = q, io.qd = qd,
    io.pf_w = pf_w, io.controls = controls, io.states = states, io.tau = tau, io.status = status, io.iters = iters,
    io.fric_active = fric_active, io.resid = resid, io.ws_mask = ws_mask, io.do_lowlevel = 1;
    std::vector<double> w1(LaneRec<30, 1>::total), w2(LaneRec<30, 2>::total), ps(LaneRec<30, 2>::smem_doubles);
    for (int i = 0; i < n; ++i) {
        int S = 0;
        for (int k = 0; k < 2 * d.h; ++k) S += contact[(size_t)i * 2 * d.h + k] ? 1 : 0;
        if (d.h == 10) {
            if (S <= 10) run_one<10, 1>(d, w1, ps, io, i, g_f32 != 0);
            else run_one<10, 2>(d, w2, ps, io, i, g_f32 != 0);
        } else {
            if (S <= 30) run_one<30, 1>(d, w1, ps, io, i, g_f32 != 0);
            else run_one<30, 2>(d, w2, ps, io, i, g_f32 != 0);
        }
    }
    return 0;
}
