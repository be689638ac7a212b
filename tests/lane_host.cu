// Host build of the lane-per-robot solver (biped_mpc_py_b200/csrc/bmpc_lane.cuh) for CPU unit tests: the SAME source
// the GPU kernel runs, one robot per call, checked against the oracle in tests/test_lane_host.py.
// Test infrastructure only: nothing in the product loads this library.
#include <vector>

#include "../biped_mpc_py_b200/csrc/bmpc_presolve.h"
#include "../biped_mpc_py_b200/csrc/bmpc_lane.cuh"

using namespace bmpc;

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
    std::vector<double> w1(LaneL<30, 1, 5>::total), w2(LaneL<30, 2, 5>::total);
    for (int i = 0; i < n; ++i) {
        int S = 0;
        for (int k = 0; k < 2 * d.h; ++k) S += contact[(size_t)i * 2 * d.h + k] ? 1 : 0;
        if (d.h == 10) {
            if (S <= 10) LaneSolver<10, 1, 5>(d, SV{w1.data()}).run(io, i);
            else LaneSolver<10, 2, 5>(d, SV{w2.data()}).run(io, i);
        } else {
            if (S <= 30) LaneSolver<30, 1, 5>(d, SV{w1.data()}).run(io, i);
            else LaneSolver<30, 2, 5>(d, SV{w2.data()}).run(io, i);
        }
    }
    return 0;
}
