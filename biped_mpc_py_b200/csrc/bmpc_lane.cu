// Translation unit of the lane-per-robot kernels (bmpc_lane.cuh): kept apart from bmpc.cu so that the kernel families
// compile in parallel.  bmpc.cu reaches the kernels through the small table below.
// BMPC_LANE_UNIT selects which horizon this unit instantiates (10, 30, or undefined = both), so that the build can split
// the instantiations over several compiler processes.
#include "bmpc_lane.cuh"
#include "bmpc_lane_api.h"

namespace bmpc {

template <int HZ, int NF>
static LaneKernelInfo pick(unsigned rowmask) {
    LaneKernel fn = lane_tick_kernel<HZ, NF, 0u>;
    if (rowmask == kRowsRef) fn = lane_tick_kernel<HZ, NF, kRowsRef>;
    else if (rowmask == kRowsSym) fn = lane_tick_kernel<HZ, NF, kRowsSym>;
    return {fn, LaneRec<HZ, NF>::total, LaneRec<HZ, NF>::smem_doubles, LaneRec<HZ, NF>::defer_floats, LaneRec<HZ, NF>::ipm_floats};
}

#if !defined(BMPC_LANE_UNIT) || BMPC_LANE_UNIT == 10
LaneKernelInfo lane_kernel_info_h10(int nf, unsigned rowmask) { return nf == 1 ? pick<10, 1>(rowmask) : pick<10, 2>(rowmask); }
#endif
#if !defined(BMPC_LANE_UNIT) || BMPC_LANE_UNIT == 30
LaneKernelInfo lane_kernel_info_h30(int nf, unsigned rowmask) { return nf == 1 ? pick<30, 1>(rowmask) : pick<30, 2>(rowmask); }
#endif

}  // namespace bmpc
