// Host-side handle on the lane-per-robot kernels compiled in bmpc_lane.cu.
#pragma once
#include "bmpc_kernels.cuh"

namespace bmpc {

// Second passes.  A warp (a CTA in lockstep) runs as many interior-point iterations and polish rounds as its slowest robot,
// with the finished lanes idle: the iteration counts inside a CTA spread from 7 to 13 and nine robots in ten are certified by
// the first polish round while the others need one to three more.  So the first pass (mode 0) PARKS the stragglers:
//   * a robot whose interior point has not converged after `ipm_inline` iterations: the iterate (u, s, lam as float) goes to a
//     lane-interleaved record of the IPM store, its index to the IPM list;
//   * a robot that needs more than `inline_rounds` polish rounds: active-row masks and multipliers go to the polish store.
// Record q of a store: element e at buf[((q / 32) * per_robot + e) * 32 + q % 32].  The later passes take the parked robots
// 32 to a warp again, rebuild the problem data from the inputs and continue: mode 2 the interior point (from the parked
// interior point, residuals recomputed) followed by the first polish round(s) - parking for the polish like the first pass -,
// then mode 1 the polish.  buf == nullptr (or a class below min_count): everything in one pass.
struct LaneDefer {
    float* buf;       // polish store
    int* list;
    int* count;       // parked robots so far (may run past cap: those robots simply carried on in the pass they were in)
    int cap;
    int inline_rounds;
    int min_count;    // classes smaller than this run everything in the first pass (a batch that is one wave of slices gains nothing
                      // from a second launch: its time is the latency of one slice)
    float* ipm_buf;   // interior-point store (nullptr: no interior-point parking)
    int* ipm_list;
    int* ipm_count;
    int ipm_cap;
    int ipm_inline;
};

typedef void (*LaneKernel)(const DevParams, const IoPtrs, const int*, const int*, int*, double*, int, const LaneDefer, int mode);

struct LaneKernelInfo {
    LaneKernel fn;        // null: no such instantiation
    int ws_doubles;       // global workspace per robot (doubles); a warp owns 32 of them, lane-interleaved
    int smem_doubles;     // dynamic shared memory per robot (doubles)
    int defer_floats;     // parked-robot record of the polish store (floats)
    int ipm_floats;       // ... of the interior-point store
};

// nf = 1 | 2 stance feet per stage; rowmask = the presolve's surviving candidate rows (bit order of bmpc_lane.cuh): the
// instantiation specialised for that row set if there is one, else the generic one
LaneKernelInfo lane_kernel_info_h10(int nf, unsigned rowmask);
LaneKernelInfo lane_kernel_info_h30(int nf, unsigned rowmask);

inline unsigned lane_rowmask(const DevParams& d) {
    unsigned m = 0u;
    for (int k = 0; k < d.mb; ++k) {
        const int kind = d.row_kind[k], arg = d.row_arg[k];
        m |= 1u << (kind == ROW_LO ? arg : kind == ROW_HI ? 5 + arg : kind == ROW_FRIC ? 10 + arg : 14 + arg);
    }
    return m;
}

}  // namespace bmpc
