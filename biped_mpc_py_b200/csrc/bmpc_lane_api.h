// Host-side handle on the lane-per-robot kernels compiled in bmpc_lane.cu.
#pragma once
#include "bmpc_kernels.cuh"

namespace bmpc {

typedef void (*LaneKernel)(const DevParams, const IoPtrs, const int*, const int*, int*, double*, int);

struct LaneKernelInfo {
    LaneKernel fn;        // null: no such instantiation
    int ws_doubles;       // global workspace per robot (doubles); a warp owns 32 of them, lane-interleaved
    int smem_doubles;     // dynamic shared memory per robot (doubles)
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
