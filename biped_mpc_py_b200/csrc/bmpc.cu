// C ABI of the B200-native batched biped MPC (see include/biped_mpc_b200.h).
// Host side: parameter presolve, workspace, kernel dispatch.  No torch types, no CPU
// fallback: every entry point either launches the CUDA kernels or returns an error.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "../../include/biped_mpc_b200.h"
#include "bmpc_lane_api.h"
#include "bmpc_presolve.h"
#include "bmpc_rollout.cuh"
#include "bmpc_small.cuh"
#include "bmpc_tick.cuh"

using namespace bmpc;

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}
#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// Every entry point works on the handle's device and leaves the caller's current device as it found it (a multi-GPU host
// process may call a handle of device k from a thread whose current device is another one).
struct DeviceGuard {
    int prev = -1, dev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int d) : dev(d) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_DEVICE(d)          \
    DeviceGuard _guard(d);    \
    CUDA_TRY(_guard.err)

typedef void (*TickKernel)(const DevParams, const IoPtrs, const int*, const int*, double*);

// lane-per-robot front end of one class (bmpc_lane.cuh): what it certifies is done, the rest goes to the class's Variant
struct LaneVariant {
    LaneKernel fn = nullptr;
    int grid = 0;
    int threads = 0;         // 32 robots per warp
    size_t smem = 0;         // dynamic shared memory per CTA: the packed cost-to-go of its robots
    size_t ws_doubles = 0;
    int min_count = 0;       // classes smaller than this stay on the warp-per-robot kernel (decided on the device)
    double* d_ws = nullptr;  // [warps][blocks][record][32] lane-interleaved block records
    // second polish pass (LaneDefer, bmpc_lane_api.h): store of the parked robots
    int defer_floats = 0, defer_cap = 0, ipm_floats = 0, ipm_cap = 0;
    float* d_defer = nullptr;
    int* d_defer_list = nullptr;
    float* d_ipm = nullptr;
    int* d_ipm_list = nullptr;
};

struct Variant {
    TickKernel fn = nullptr;
    size_t smem = 0;
    int threads = 0;
    int per_cta = 1;   // robots (thread groups) per CTA
    int resident = 0;  // thread groups that fit on the device at once
    size_t scratch_doubles = 0;  // per resident CTA: one copy of the tile matrix H
    double* d_scratch = nullptr;
};

}  // namespace

struct bmpc_handle {
    int device = 0;
    int max_batch = 0;
    int num_sms = 0;
    bmpc_params host_params;
    DevParams dp;
    Variant bucket[2];
    Variant lowlat;           // walking class for small batches: 128 threads per robot (latency, not throughput), or empty
    LaneVariant lane[2];      // LB = 5, mx pinned: lane-per-robot front end per class (batches >= lane_min), or empty
    int lane_min = 0;
    // tunables (bmpc_set_option); -1 = the built-in default
    int opt_lane_mode = -1;        // 2 both classes, 1 walking class only, 0 warp-per-robot kernels only
    int opt_lane_min = -1;         // overrides the three size gates
    int opt_lane_ctas_per_sm = -1; // resident CTAs per SM of the lane kernels
    int opt_lane_ctas_standing = -1;  // ... of the standing class only (-1: as the walking class)
    int opt_lane_warps = -1;       // warps (32 robots each) per CTA
    int opt_lowlat = -1;           // 0 disables the 128-thread walking variant for batches <= 8
    int opt_lane_ipm_inline = -1;     // interior-point iterations of the first lane pass before a robot is parked (0: never)
    int opt_lane_defer_cap = -1;      // size of the two park stores in robots (-1: a quarter of max_batch, at least 4,096)
    int opt_lane_defer_min = -1;      // smallest class that uses the second pass (-1: two waves of slices)
    int opt_lane_inline_rounds = -1;  // polish rounds of the first lane pass before a robot is parked for the second (0: one pass)
    Variant fallback;         // dense re-solve of instances the stage-wise class-1 kernel did not certify (h = 30), or empty
    int* d_lists = nullptr;   // [5][max_batch]: two classes, the h = 30 fallback list, the two lane-residual lists
    int* d_counts = nullptr;  // [32]: list count i has its dynamic work counter at i + 3 (lists 0-2 and 6-7); 12, 13 lane slice counters; 20, 21 parked
                              //       robots (polish) of the lane kernels
                              //       with the polish pass's slice counters at 22, 23; 24, 25 parked robots (interior point), slice counters at 26, 27;
                              //       16, 17 last-resort list counts with their slice counters at 18, 19
    int64_t launches = 0;
    // closed-loop rollout workspace (allocated on the first bmpc_rollout call, max_batch sized)
    struct {
        uint8_t* contact = nullptr;
        int32_t* phase_k = nullptr;
        double* t_swing = nullptr;
        double* controls = nullptr;
        double* tau = nullptr;
        int32_t* status = nullptr;
        int32_t* iters = nullptr;
        int32_t* ws_mask = nullptr;
    } ro;
    int warm_enabled = 0;        // bmpc_warm_start: bmpc_step / bmpc_solve start from the previous call's active set
    int warm_valid = 0;          // the store holds the masks of a previous call ...
    int warm_n = 0;              // ... for its first warm_n robots
    int timing = 0;              // record CUDA events around each kernel of a tick
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // the two lane-per-robot class kernels run concurrently: the standing class on the caller's stream, the walking class on
    // this side stream, so that the classes share one tail instead of each paying its own
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};

namespace {

int build_dev_params(const bmpc_params& P, DevParams& d) {
    std::string err;
    if (build_dev_params_impl(P, d, err)) return fail(err);
    return 0;
}

template <int HZ, int SMAX, int LB, int NT, int NW, bool MG = false, bool RIC = false>
int setup_variant(Variant& v, int num_sms, int mb) {
    using L = TickLayout<HZ, SMAX, LB, MG, RIC>;
    v.fn = mpc_tick2_kernel<HZ, SMAX, LB, NT, NW, MG, RIC>;
    v.smem = L::bytes(mb) * NW + 16;  // + the CTA-wide lockstep mbarrier
#ifdef BMPC_EXPERIMENTS
    if (const char* ep = getenv("BMPC_PAD_SMEM")) v.smem += (size_t)atoi(ep);  // occupancy experiments only
#endif
    v.threads = NT * NW;
    v.per_cta = NW;
    v.scratch_doubles = L::g_total;
    CUDA_TRY(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.fn, NT * NW, v.smem));
    if (per_sm < 1) return fail("kernel does not fit on an SM");
    v.resident = per_sm * num_sms * NW;
    CUDA_TRY(cudaMalloc(&v.d_scratch, sizeof(double) * v.scratch_doubles * (size_t)v.resident));
    return 0;
}

void free_lane(LaneVariant& v) {
    cudaFree(v.d_ws), cudaFree(v.d_defer), cudaFree(v.d_defer_list), cudaFree(v.d_ipm), cudaFree(v.d_ipm_list);
    v.d_ws = nullptr, v.d_defer = nullptr, v.d_defer_list = nullptr, v.d_ipm = nullptr, v.d_ipm_list = nullptr;
}

int setup_lane(LaneVariant& v, const DevParams& d, int nf, int num_sms, int ctas_per_sm, int warps, int max_batch, int cap_override) {
    const LaneKernelInfo k = d.h == 30 ? lane_kernel_info_h30(nf, lane_rowmask(d)) : lane_kernel_info_h10(nf, lane_rowmask(d));
    if (!k.fn) return fail("no lane kernel for this horizon");
    free_lane(v);
    v.defer_floats = k.defer_floats;
    v.defer_cap = cap_override > 0 ? (cap_override + 31) / 32 * 32
                                   : std::max(4096, ((max_batch + 3) / 4 + 31) / 32 * 32);  // a quarter of the batch can be parked for the polish (about one robot in eight is)
    v.ipm_floats = k.ipm_floats;
    v.ipm_cap = v.defer_cap;                                              // and for the interior point (about one robot in forty is)
    v.fn = k.fn;
    v.threads = 32 * warps;
    v.smem = sizeof(double) * (size_t)k.smem_doubles * 32 * warps;
    CUDA_TRY(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.fn, v.threads, v.smem));
    if (per_sm < 1) return fail("lane kernel does not fit on an SM");
    if (ctas_per_sm > 0) per_sm = std::min(per_sm, ctas_per_sm);
    v.grid = std::min(per_sm * num_sms, (max_batch + v.threads - 1) / v.threads);
    v.ws_doubles = (size_t)k.ws_doubles * 32 * warps * (size_t)v.grid;  // allocated on first use (throughput batches only)
    return 0;
}

// warm-start store: every entry -1 (= no guess for this block) until a tick has written it
int alloc_ws_mask(bmpc_handle* h, cudaStream_t st) {
    if (h->ro.ws_mask) return 0;
    const size_t bytes = (size_t)h->max_batch * h->dp.h * 2 * sizeof(int32_t);
    CUDA_TRY(cudaMalloc(&h->ro.ws_mask, bytes));
    CUDA_TRY(cudaMemsetAsync(h->ro.ws_mask, 0xFF, bytes, st));
    return 0;
}

int launch_tick(bmpc_handle* h, int n, IoPtrs io, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n > h->max_batch) return fail("batch larger than max_batch given to bmpc_create");
    auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const bool handle_warm = h->warm_enabled && io.ws_mask == nullptr;  // handle-level warm start (bmpc_warm_start); bmpc_rollout manages its own
    if (handle_warm) {
        if (alloc_ws_mask(h, st)) return 1;
        io.ws_mask = h->ro.ws_mask;
        // the store holds the masks of the first warm_n robots of the last call that went through: a larger batch starts cold
        io.warm = (h->warm_valid && n <= h->warm_n) ? 1 : 0;
        h->warm_valid = 0;  // (set again below, once every launch of this call has been accepted)
    }
    io.use_tma = aligned16(io.x_fb) && aligned16(io.foot) &&
                 (!io.do_lowlevel || (aligned16(io.q) && aligned16(io.qd) && aligned16(io.pf_w)));
    CUDA_TRY(cudaMemsetAsync(h->d_counts, 0, 32 * sizeof(int), st));  // list counts + dynamic work counters
    if (h->timing) CUDA_TRY(cudaEventRecord(h->ev[0], st));
    classify_kernel<<<(n + 255) / 256, 256, 0, st>>>(io.contact, n, h->dp.h, h->max_batch, h->d_lists, h->d_counts);
    if (h->timing) CUDA_TRY(cudaEventRecord(h->ev[1], st));
    // Each class is an independent chain:  [lane-per-robot kernel] -> collect what it did not certify -> warp-per-robot
    // kernel (-> h = 30: dense re-solve of the rest).  Throughput batches run the two chains CONCURRENTLY - the standing
    // class (twice the work per robot, usually the smaller class) on the caller's stream, the walking class on a side
    // stream - so that the classes share one tail instead of each paying its own.  With timing enabled the kernels run
    // one after the other, each bracketed by events, so that the per-kernel durations are not blurred by the overlap.
    const bool use_lane = n >= h->lane_min && !io.warm && (h->lane[0].fn || h->lane[1].fn);
    const bool concurrent = use_lane && !h->timing;
    if (concurrent) {
        if (!h->side) {
            CUDA_TRY(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&h->join, cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventRecord(h->fork, st));
        CUDA_TRY(cudaStreamWaitEvent(h->side, h->fork, 0));
    }
    auto lane_launch = [&](int b, cudaStream_t ls) -> int {
        LaneVariant& lv = h->lane[b];
        if (!use_lane || !lv.fn) return 0;
        // one THREAD per robot (32 robots share every instruction); a class below its size gate is left alone (decided on
        // the device: the kernel returns at once and collect_or_all_kernel passes the whole list on)
        if (!lv.d_ws) CUDA_TRY(cudaMalloc(&lv.d_ws, sizeof(double) * lv.ws_doubles));
        const int inline_rounds = h->opt_lane_inline_rounds >= 0 ? h->opt_lane_inline_rounds : 1;
        // h = 10: 9 iterations finish 86 % of the robots (mean 8.5, slowest of a 128-robot CTA 11.6); measured at 262,144 robots:
        // parking after 8 / 9 / 10 / 11 iterations or never: (store overflows) / 28.6 / 28.9 / 29.4 / 29.8 ms per tick
        const int ipm_inline = h->opt_lane_ipm_inline >= 0 ? h->opt_lane_ipm_inline : (h->dp.h == 10 ? 9 : 0);
        LaneDefer df{};
        if (inline_rounds > 0) {
            if (!lv.d_defer) CUDA_TRY(cudaMalloc(&lv.d_defer, sizeof(float) * (size_t)lv.defer_floats * lv.defer_cap));
            if (!lv.d_defer_list) CUDA_TRY(cudaMalloc(&lv.d_defer_list, sizeof(int) * (size_t)lv.defer_cap));
            // worth further launches from about two waves of slices on (measured: 32,768 robots 6.6 ms in one pass, 6.9 in two; 262,144
            // robots 33.0 and 29.7 ms)
            const int two_waves = 2 * lv.grid * lv.threads;
            df.buf = lv.d_defer, df.list = lv.d_defer_list, df.count = h->d_counts + 20 + b, df.cap = lv.defer_cap;
            df.inline_rounds = inline_rounds;
            df.min_count = h->opt_lane_defer_min >= 0 ? h->opt_lane_defer_min : two_waves;
            if (ipm_inline > 0) {
                if (!lv.d_ipm) CUDA_TRY(cudaMalloc(&lv.d_ipm, sizeof(float) * (size_t)lv.ipm_floats * lv.ipm_cap));
                if (!lv.d_ipm_list) CUDA_TRY(cudaMalloc(&lv.d_ipm_list, sizeof(int) * (size_t)lv.ipm_cap));
                df.ipm_buf = lv.d_ipm, df.ipm_list = lv.d_ipm_list, df.ipm_count = h->d_counts + 24 + b, df.ipm_cap = lv.ipm_cap;
                df.ipm_inline = ipm_inline;
            }
        }
        const int grid = std::min(lv.grid, (n + lv.threads - 1) / lv.threads);
        lv.fn<<<grid, lv.threads, lv.smem, ls>>>(h->dp, io, h->d_lists + (size_t)b * h->max_batch, h->d_counts + b, h->d_counts + 12 + b,
                                                lv.d_ws, lv.min_count, df, 0);
        h->launches += 1;
        // later passes over the robots the first pass parked, 32 to a warp again: the interior-point stragglers (which may park for
        // the polish in turn), then the polish stragglers
        if (df.ipm_buf) {
            lv.fn<<<std::min(grid, (lv.ipm_cap + lv.threads - 1) / lv.threads), lv.threads, lv.smem, ls>>>(
                h->dp, io, nullptr, nullptr, h->d_counts + 26 + b, lv.d_ws, 0, df, 2);
            h->launches += 1;
        }
        if (df.buf) {
            lv.fn<<<std::min(grid, (lv.defer_cap + lv.threads - 1) / lv.threads), lv.threads, lv.smem, ls>>>(
                h->dp, io, nullptr, nullptr, h->d_counts + 22 + b, lv.d_ws, 0, df, 1);
            h->launches += 1;
        }
        return 0;
    };
    auto warp_launch = [&](int b, cudaStream_t ls) -> int {
        // real-time use (N = 1 .. 8): every walking robot gets a whole 128-thread CTA (0.26 ms instead of 0.32 ms for one
        // robot end to end).  Same optimum, but reductions run in a different order, so results of batches <= 8 may
        // differ in the last bits from the throughput kernel.
        const Variant& v = (b == 0 && h->lowlat.fn && h->opt_lowlat != 0 && n <= 8) ? h->lowlat : h->bucket[b];
        // persistent thread groups: as many as fit on the device, each strides over its bucket's work list
        const int grid = (std::min(n, v.resident) + v.per_cta - 1) / v.per_cta;
        const int* list = h->d_lists + (size_t)b * h->max_batch;
        const int* cnt = h->d_counts + b;
        if (use_lane && h->lane[b].fn) {
            int* rlist = h->d_lists + (size_t)(3 + b) * h->max_batch;
            collect_or_all_kernel<<<(n + 255) / 256, 256, 0, ls>>>(list, cnt, io.status, rlist, h->d_counts + 6 + b, h->lane[b].min_count);
            list = rlist, cnt = h->d_counts + 6 + b;
            h->launches += 1;
        }
        v.fn<<<grid, v.threads, v.smem, ls>>>(h->dp, io, list, cnt, v.d_scratch);
        h->launches += 1;
        return 0;
    };
    if (concurrent) {
        if (lane_launch(1, st) || lane_launch(0, h->side) || warp_launch(1, st) || warp_launch(0, h->side)) return 1;
        CUDA_TRY(cudaEventRecord(h->join, h->side));
        CUDA_TRY(cudaStreamWaitEvent(st, h->join, 0));
    } else {
        if (lane_launch(0, st)) return 1;
        if (h->timing) CUDA_TRY(cudaEventRecord(h->ev[2], st));
        if (lane_launch(1, st)) return 1;
        if (h->timing) CUDA_TRY(cudaEventRecord(h->ev[3], st));
        if (warp_launch(0, st)) return 1;
        if (h->timing) CUDA_TRY(cudaEventRecord(h->ev[4], st));
        if (warp_launch(1, st)) return 1;
    }
    if (h->fallback.fn) {
        // h = 30: instances of EITHER class that the stage-wise kernels did not certify: dense re-solve (handles any S <= 2h)
        const Variant& f = h->fallback;
        for (int c = 0; c < 2; ++c)
            collect_uncertified_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->d_lists + (size_t)c * h->max_batch, h->d_counts + c, io.status,
                                                                       h->d_lists + 2 * (size_t)h->max_batch, h->d_counts + 2);
        const int fgrid = (std::min(n, f.resident) + f.per_cta - 1) / f.per_cta;
        f.fn<<<fgrid, f.threads, f.smem, st>>>(h->dp, io, h->d_lists + 2 * (size_t)h->max_batch, h->d_counts + 2, f.d_scratch);
        h->launches += 3;
    }
    if (n > 8 && (h->lane[0].fn || h->lane[1].fn)) {
        // Last resort for what is still not certified (status 1 / 2): the lane-per-robot solver with a long polish budget.  A
        // nearly degenerate instance whose first active-set guesses are far off (one in ~65,000 at h = 30; none seen at h = 10)
        // walks to the optimal set one row per block and round, which takes more rounds than the kernels above allow.
        // (Batches of 1 .. 8 robots - the latency path - skip this step.)
        DevParams last = h->dp;
        last.polish_rounds = 128;
        last.lane_sync = 0;
        for (int b = 0; b < 2; ++b) {
            LaneVariant& lv = h->lane[b];
            if (!lv.fn) continue;
            if (!lv.d_ws) CUDA_TRY(cudaMalloc(&lv.d_ws, sizeof(double) * lv.ws_doubles));
            int* rlist = h->d_lists + (size_t)(3 + b) * h->max_batch;
            collect_uncertified_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->d_lists + (size_t)b * h->max_batch, h->d_counts + b, io.status, rlist,
                                                                       h->d_counts + 16 + b);
            lv.fn<<<std::min(lv.grid, 8), lv.threads, lv.smem, st>>>(last, io, rlist, h->d_counts + 16 + b, h->d_counts + 18 + b, lv.d_ws, 1,
                                                                     LaneDefer{}, 0);
            h->launches += 2;
        }
    }
    if (h->timing) CUDA_TRY(cudaEventRecord(h->ev[5], st));
    h->launches += 1;  // classify
    CUDA_TRY(cudaGetLastError());
    if (handle_warm) h->warm_valid = 1, h->warm_n = n;
    return 0;
}

// Lane-per-robot front end (bmpc_lane.cuh) for throughput batches (n >= lane_min): one THREAD per robot, 32 robots per
// instruction; whatever it does not certify falls through to the warp-per-robot kernel of the class.  It needs the reference's
// limit structure (five free components, mx pinned: tau_max[0] = tau_min[0], MPC.py:47).
// Size gates: a 32-robot slice takes a fixed time however small the batch, so the front end pays only when a class can
// occupy the machine; the class sizes are only known on the device, so the lane kernel itself returns at once for a small
// class and collect_or_all_kernel passes the whole list on.
int setup_lanes(bmpc_handle* h) {
    for (int b = 0; b < 2; ++b) {
        free_lane(h->lane[b]);
        h->lane[b] = LaneVariant();
    }
    h->lane_min = 0;
    const DevParams& d = h->dp;
    if (d.LB != 5 || d.npinned != 1 || d.pinned[0] != 3 || d.mb > 16) return 0;
    const int mode = h->opt_lane_mode >= 0 ? h->opt_lane_mode : 2;
    if (mode == 0) return 0;
    const int warps = h->opt_lane_warps > 0 ? h->opt_lane_warps : 4;
    const int ctas = h->opt_lane_ctas_per_sm > 0 ? h->opt_lane_ctas_per_sm : 0;
    const bool h30 = d.h == 30;
    // measured crossovers against the warp-per-robot kernels (tools/gate_probe.py, profiles/r2_summary.md): h = 10 walking
    // class ~8 k robots (3.8 vs 3.6 ms), standing ~4 k (6.2 vs 6.2 ms); h = 30 walking ~2 k (17 vs 18 ms), standing ~1 k (40 vs 45 ms)
    h->lane_min = h->opt_lane_min >= 0 ? h->opt_lane_min : (h30 ? 1024 : 4096);
    h->lane[0].min_count = h->opt_lane_min >= 0 ? h->opt_lane_min : (h30 ? 2048 : 8192);
    h->lane[1].min_count = h->opt_lane_min >= 0 ? h->opt_lane_min : (h30 ? 1024 : 4096);
    int rc = setup_lane(h->lane[0], d, 1, h->num_sms, ctas, warps, h->max_batch, h->opt_lane_defer_cap);
    if (!rc && mode >= 2) {
        const int mc = h->lane[1].min_count;
        rc = setup_lane(h->lane[1], d, 2, h->num_sms, h->opt_lane_ctas_standing > 0 ? h->opt_lane_ctas_standing : ctas, warps, h->max_batch,
                        h->opt_lane_defer_cap);
        h->lane[1].min_count = mc;
    }
    return rc;
}

}  // namespace

extern "C" {

const char* bmpc_last_error(void) { return g_err.c_str(); }
int bmpc_abi_version(void) { return BMPC_ABI_VERSION; }

int bmpc_create(const bmpc_params* params, int device, int max_batch, bmpc_handle** out) {
    if (!params || !out || max_batch <= 0) return fail("bmpc_create: bad arguments");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("bmpc_create: device index out of range");
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("this build targets sm_100a (Blackwell B200) only");
    bmpc_handle* h = new bmpc_handle();
    h->device = device;
    h->max_batch = max_batch;
    h->num_sms = prop.multiProcessorCount;
    h->host_params = *params;
    if (build_dev_params(*params, h->dp)) {
        delete h;
        return 1;
    }
    int rc = 0;
    // The shipped configuration of every class.  Building with -DBMPC_EXPERIMENTS adds the alternative instantiations the
    // measurements in profiles/r1_summary.md were taken with (BMPC_NW_WALK = 1 | 5, BMPC_NT_STAND = 256, BMPC_H30 = dense | ric,
    // BMPC_RIC_NT = 32), selected through environment variables.
    const int sms = h->num_sms, mb = h->dp.mb;
#ifdef BMPC_EXPERIMENTS
    auto envi = [](const char* k, int dflt) { const char* v = getenv(k); return v ? atoi(v) : dflt; };
    const int nww = envi("BMPC_NW_WALK", 8), nts = envi("BMPC_NT_STAND", 128), rnt = envi("BMPC_RIC_NT", 128);
    const char* eh = getenv("BMPC_H30");
    const std::string h30mode = eh ? eh : "ric";
#endif
    if (h->dp.h == 30) {
        // h = 30 (BASELINE.json configs[3]).  Both classes use the stage-wise Riccati backend (110 ms vs 135 ms dense for the
        // walking class, 59 ms vs 223 ms for the standing class at 16,384 instances), followed by a dense re-solve (tile
        // matrix in the L2 scratch) of the few instances it does not certify.
        if (h->dp.LB == 6) {  // no pinned component: same three kernels with 6x6 tiles / 12 inputs per stage
            rc = setup_variant<30, 30, 6, 128, 1, false, true>(h->bucket[0], sms, mb) ||
                 setup_variant<30, 60, 6, 128, 1, false, true>(h->bucket[1], sms, mb) ||
                 setup_variant<30, 60, 6, 256, 1, true>(h->fallback, sms, mb);
        } else {
#ifdef BMPC_EXPERIMENTS
            if (h30mode == "hybrid")
                rc = setup_variant<30, 30, 5, 256, 1>(h->bucket[0], sms, mb) ||
                     setup_variant<30, 60, 5, 128, 1, false, true>(h->bucket[1], sms, mb) ||
                     setup_variant<30, 60, 5, 256, 1, true>(h->fallback, sms, mb);
            else if (h30mode == "dense")
                rc = setup_variant<30, 30, 5, 256, 1>(h->bucket[0], sms, mb) || setup_variant<30, 60, 5, 256, 1, true>(h->bucket[1], sms, mb);
            else if (h30mode == "ric")
                rc = (rnt == 32 ? (setup_variant<30, 30, 5, 32, 1, false, true>(h->bucket[0], sms, mb) ||
                                   setup_variant<30, 60, 5, 32, 1, false, true>(h->bucket[1], sms, mb))
                                : (setup_variant<30, 30, 5, 128, 1, false, true>(h->bucket[0], sms, mb) ||
                                   setup_variant<30, 60, 5, 128, 1, false, true>(h->bucket[1], sms, mb))) ||
                     setup_variant<30, 60, 5, 256, 1, true>(h->fallback, sms, mb);
            else
#endif
                rc = setup_variant<30, 30, 5, 128, 1, false, true>(h->bucket[0], sms, mb) ||
                     setup_variant<30, 60, 5, 128, 1, false, true>(h->bucket[1], sms, mb) ||
                     setup_variant<30, 60, 5, 256, 1, true>(h->fallback, sms, mb);
            if (!rc) rc = setup_lanes(h);
        }
    } else if (h->dp.LB == 5) {
#ifdef BMPC_EXPERIMENTS
        if (nww == 1) rc = setup_variant<10, 10, 5, 32, 1>(h->bucket[0], sms, mb);
        else if (nww == 5) rc = setup_variant<10, 10, 5, 32, 5>(h->bucket[0], sms, mb);
        else
#endif
            rc = setup_variant<10, 10, 5, 32, 8>(h->bucket[0], sms, mb);  // walking: 8 robots per CTA, one warp each
#ifdef BMPC_EXPERIMENTS
        if (!rc && nts == 256) rc = setup_variant<10, 20, 5, 256, 1>(h->bucket[1], sms, mb);
        else
#endif
            if (!rc) rc = setup_variant<10, 20, 5, 128, 1>(h->bucket[1], sms, mb);  // standing: one 128-thread CTA per robot
        if (!rc) rc = setup_lanes(h);
        if (!rc) rc = setup_variant<10, 10, 5, 128, 1>(h->lowlat, sms, mb);  // batches <= 8: 128 threads per walking robot
    } else {
        rc = setup_variant<10, 10, 6, 32, 6>(h->bucket[0], sms, mb) || setup_variant<10, 20, 6, 128, 1>(h->bucket[1], sms, mb);
    }
    if (rc) {
        bmpc_destroy(h);  // frees whatever the variants before the failing one allocated
        return 1;
    }
    e = cudaMalloc(&h->d_lists, sizeof(int) * 5 * (size_t)max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_counts, sizeof(int) * 32);
    if (e != cudaSuccess) {
        bmpc_destroy(h);
        return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    *out = h;
    return 0;
}

int bmpc_destroy(bmpc_handle* h) {
    if (!h) return 0;
    DeviceGuard _guard(h->device);
    cudaFree(h->d_lists);
    cudaFree(h->d_counts);
    for (int b = 0; b < 2; ++b) cudaFree(h->bucket[b].d_scratch);
    cudaFree(h->fallback.d_scratch);
    cudaFree(h->lowlat.d_scratch);
    free_lane(h->lane[0]), free_lane(h->lane[1]);
    cudaFree(h->ro.contact), cudaFree(h->ro.phase_k), cudaFree(h->ro.t_swing), cudaFree(h->ro.controls);
    cudaFree(h->ro.tau), cudaFree(h->ro.status), cudaFree(h->ro.iters), cudaFree(h->ro.ws_mask);
    for (int i = 0; i < 6; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->fork) cudaEventDestroy(h->fork);
    if (h->join) cudaEventDestroy(h->join);
    if (h->side) cudaStreamDestroy(h->side);
    delete h;
    return 0;
}

int bmpc_step(bmpc_handle* h, int n, const double* x_fb, const int32_t* phase_k, const double* t_swing,
              const double* foot, const uint8_t* contact, const double* q, const double* qd, const double* pf_w,
              double* controls, double* states, double* tau, int32_t* status, int32_t* iters, uint8_t* fric_active,
              double* resid, void* stream) {
    if (!h) return fail("bmpc_step: null handle");
    if (!x_fb || !phase_k || !t_swing || !foot || !contact || !q || !qd || !pf_w || !controls || !tau || !status ||
        !iters)
        return fail("bmpc_step: null required pointer");
    ON_DEVICE(h->device);
    IoPtrs io;
    memset(&io, 0, sizeof(io));
    io.x_fb = x_fb, io.phase_k = phase_k, io.t_swing = t_swing, io.foot = foot, io.contact = contact;
    io.q = q, io.qd = qd, io.pf_w = pf_w;
    io.controls = controls, io.states = states, io.tau = tau, io.status = status, io.iters = iters;
    io.fric_active = fric_active, io.resid = resid;
    io.do_lowlevel = 1;
    return launch_tick(h, n, io, static_cast<cudaStream_t>(stream));
}

int bmpc_solve(bmpc_handle* h, int n, const double* x_fb, const int32_t* phase_k, const double* foot,
               const uint8_t* contact, double* controls, double* states, int32_t* status, int32_t* iters,
               uint8_t* fric_active, double* resid, void* stream) {
    if (!h) return fail("bmpc_solve: null handle");
    if (!x_fb || !phase_k || !foot || !contact || !controls || !status || !iters)
        return fail("bmpc_solve: null required pointer");
    ON_DEVICE(h->device);
    IoPtrs io;
    memset(&io, 0, sizeof(io));
    io.x_fb = x_fb, io.phase_k = phase_k, io.foot = foot, io.contact = contact;
    io.controls = controls, io.states = states, io.status = status, io.iters = iters;
    io.fric_active = fric_active, io.resid = resid;
    io.do_lowlevel = 0;
    return launch_tick(h, n, io, static_cast<cudaStream_t>(stream));
}

int bmpc_lowlevel(bmpc_handle* h, int n, const double* x_fb, const double* t_swing, const double* pf_w,
                  const double* q, const double* qd, const uint8_t* contact0, const double* u0, double* tau,
                  void* stream) {
    if (!h) return fail("bmpc_lowlevel: null handle");
    if (!x_fb || !t_swing || !pf_w || !q || !qd || !contact0 || !u0 || !tau)
        return fail("bmpc_lowlevel: null required pointer");
    if (n <= 0) return 0;
    ON_DEVICE(h->device);
    const int threads = 128, total = 2 * n;
    lowlevel_kernel<<<(total + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        h->dp, n, x_fb, t_swing, pf_w, q, qd, contact0, u0, tau);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int bmpc_foot_positions(bmpc_handle* h, int n, const double* x_fb, const double* q, double* pf_w, void* stream) {
    if (!h) return fail("bmpc_foot_positions: null handle");
    if (!x_fb || !q || !pf_w) return fail("bmpc_foot_positions: null required pointer");
    if (n <= 0) return 0;
    ON_DEVICE(h->device);
    const int threads = 128, total = 2 * n;
    foot_positions_kernel<<<(total + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        h->dp, n, x_fb, q, pf_w);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int bmpc_rollout(bmpc_handle* h, int n, int ticks, double* x, double* foot, int32_t* tick, const uint8_t* gait,
                 const double* q, const double* qd, int warm_start, int n_log, double* x_log, double* foot_log,
                 double* u0_log, double* tau_log, uint64_t* stats, void* stream) {
    if (!h) return fail("bmpc_rollout: null handle");
    if (!x || !foot || !tick || !gait || !q || !qd) return fail("bmpc_rollout: null required pointer");
    if (n <= 0 || ticks <= 0) return 0;
    if (n > h->max_batch) return fail("batch larger than max_batch given to bmpc_create");
    if (n_log < 0 || n_log > n) return fail("bmpc_rollout: n_log must be in [0, n]");
    if (n_log > 0 && (!x_log || !foot_log || !u0_log || !tau_log)) return fail("bmpc_rollout: n_log > 0 needs all four log buffers");
    ON_DEVICE(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int hz = h->dp.h;
    const size_t nb = (size_t)h->max_batch;
    h->warm_valid = 0;  // the rollout shares the warm-start store: a later warm bmpc_step starts cold
    if (alloc_ws_mask(h, st)) return 1;
    // (each buffer on its own: a failed allocation leaves the others in place for the next call and for bmpc_destroy)
    if (!h->ro.contact) CUDA_TRY(cudaMalloc(&h->ro.contact, nb * hz * 2));
    if (!h->ro.phase_k) CUDA_TRY(cudaMalloc(&h->ro.phase_k, nb * sizeof(int32_t)));
    if (!h->ro.t_swing) CUDA_TRY(cudaMalloc(&h->ro.t_swing, nb * sizeof(double)));
    if (!h->ro.controls) CUDA_TRY(cudaMalloc(&h->ro.controls, nb * hz * 12 * sizeof(double)));
    if (!h->ro.tau) CUDA_TRY(cudaMalloc(&h->ro.tau, nb * 10 * sizeof(double)));
    if (!h->ro.status) CUDA_TRY(cudaMalloc(&h->ro.status, nb * sizeof(int32_t)));
    if (!h->ro.iters) CUDA_TRY(cudaMalloc(&h->ro.iters, nb * sizeof(int32_t)));
    RolloutPtrs r;
    memset(&r, 0, sizeof(r));
    r.x = x, r.foot = foot, r.tick = tick, r.gait = gait, r.q = q;
    r.contact = h->ro.contact, r.phase_k = h->ro.phase_k, r.t_swing = h->ro.t_swing;
    r.controls = h->ro.controls, r.tau = h->ro.tau, r.status = h->ro.status, r.iters = h->ro.iters;
    r.n_log = n_log, r.x_log = x_log, r.foot_log = foot_log, r.u0_log = u0_log, r.tau_log = tau_log;
    r.stats = reinterpret_cast<unsigned long long*>(stats);
    IoPtrs io;
    memset(&io, 0, sizeof(io));
    io.x_fb = x, io.phase_k = h->ro.phase_k, io.t_swing = h->ro.t_swing, io.foot = foot, io.contact = h->ro.contact;
    io.q = q, io.qd = qd, io.pf_w = foot;  // R3: pf_w = foot
    io.controls = h->ro.controls, io.tau = h->ro.tau, io.status = h->ro.status, io.iters = h->ro.iters;
    io.do_lowlevel = 1;
    io.ws_mask = warm_start ? h->ro.ws_mask : nullptr;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    const int saved_warm = h->warm_enabled;
    h->warm_enabled = 0;  // the rollout passes its own warm-start store (or none) to the tick
    for (int k = 0; k < ticks; ++k) {
        rollout_prepare_kernel<<<blocks, threads, 0, st>>>(h->dp, r, n, k);
        io.warm = (warm_start && k > 0) ? 1 : 0;  // the first tick of a call is always cold
        if (launch_tick(h, n, io, st)) {
            h->warm_enabled = saved_warm;
            return 1;
        }
        rollout_advance_kernel<<<blocks, threads, 0, st>>>(h->dp, r, n, k);
        h->launches += 2;
    }
    h->warm_enabled = saved_warm;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int bmpc_warm_start(bmpc_handle* h, int mode) {
    if (!h) return fail("bmpc_warm_start: null handle");
    if (mode < 0 || mode > 2) return fail("bmpc_warm_start: mode must be 0 (off), 1 (on) or 2 (forget the stored active sets)");
    if (mode == 2) {
        h->warm_valid = 0;
        return 0;
    }
    h->warm_enabled = mode;
    h->warm_valid = 0;
    return 0;
}

int bmpc_debug_assemble(bmpc_handle* h, const double* x_fb, const int32_t* phase_k, const double* foot,
                        const uint8_t* contact, double* Hc_out, double* g_out, int32_t* n_out, void* stream) {
    if (!h) return fail("bmpc_debug_assemble: null handle");
    if (!x_fb || !phase_k || !foot || !contact || !Hc_out || !g_out || !n_out)
        return fail("bmpc_debug_assemble: null required pointer");
    ON_DEVICE(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* scratch = nullptr;  // controls + status + iters for the single instance
    const int hz = h->dp.h;
    CUDA_TRY(cudaMalloc(&scratch, sizeof(double) * (hz * 12 + 4)));
    IoPtrs io;
    memset(&io, 0, sizeof(io));
    io.x_fb = x_fb, io.phase_k = phase_k, io.foot = foot, io.contact = contact;
    io.controls = scratch;
    io.status = reinterpret_cast<int32_t*>(scratch + hz * 12);
    io.iters = io.status + 1;
    io.dbg_H = Hc_out, io.dbg_g = g_out, io.dbg_n = n_out;
    const int saved_warm = h->warm_enabled;
    h->warm_enabled = 0;
    int rc = launch_tick(h, 1, io, st);
    h->warm_enabled = saved_warm;
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(scratch);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(std::string("debug kernel: ") + cudaGetErrorString(e));
    return 0;
}

int bmpc_set_option(bmpc_handle* h, const char* name, int value) {
    if (!h || !name) return fail("bmpc_set_option: null argument");
    ON_DEVICE(h->device);
    const std::string k(name);
    if (k == "lane_mode") h->opt_lane_mode = value;
    else if (k == "lane_min") h->opt_lane_min = value;
    else if (k == "lane_ctas_per_sm") h->opt_lane_ctas_per_sm = value;
    else if (k == "lane_ctas_standing") h->opt_lane_ctas_standing = value;
    else if (k == "lane_warps") h->opt_lane_warps = value;
    else if (k == "lowlat") { h->opt_lowlat = value; return 0; }
    else if (k == "lane_inline_rounds") { h->opt_lane_inline_rounds = value; return 0; }
    else if (k == "lane_defer_min") { h->opt_lane_defer_min = value; return 0; }
    else if (k == "lane_defer_cap") h->opt_lane_defer_cap = value;
    else if (k == "lane_ipm_inline") { h->opt_lane_ipm_inline = value; return 0; }
    else if (k == "lane_prefetch") { h->dp.lane_prefetch = value != 0; return 0; }
    else if (k == "lane_sync") { h->dp.lane_sync = value; return 0; }
    else if (k == "polish_rounds") { if (value < 1) return fail("bmpc_set_option: polish_rounds must be >= 1"); h->dp.polish_rounds = value; return 0; }
    else return fail("bmpc_set_option: unknown option '" + k + "'");
    CUDA_TRY(cudaDeviceSynchronize());  // the workspace of the lane kernels is re-sized
    return setup_lanes(h);
}

int64_t bmpc_launch_count(const bmpc_handle* h) { return h ? h->launches : 0; }

int bmpc_enable_timing(bmpc_handle* h, int enable) {
    if (!h) return fail("bmpc_enable_timing: null handle");
    ON_DEVICE(h->device);
    if (enable && !h->ev[0])
        for (int i = 0; i < 6; ++i) CUDA_TRY(cudaEventCreate(&h->ev[i]));
    h->timing = enable ? 1 : 0;
    return 0;
}

int bmpc_last_timing(bmpc_handle* h, float* ms5) {
    if (!h || !ms5 || !h->ev[0]) return fail("bmpc_last_timing: timing was not enabled");
    ON_DEVICE(h->device);
    CUDA_TRY(cudaEventSynchronize(h->ev[5]));
    for (int i = 0; i < 5; ++i) CUDA_TRY(cudaEventElapsedTime(&ms5[i], h->ev[i], h->ev[i + 1]));
    return 0;
}

int bmpc_measure_fma_peak(int device, int fp64, double* tflops_out) {
    if (!tflops_out) return fail("bmpc_measure_fma_peak: null output");
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
    struct Scratch {  // released on every return path
        void* buf = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Scratch() {
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
            cudaFree(buf);
        }
    } sc;
    void*& buf = sc.buf;
    cudaEvent_t &e0 = sc.e0, &e1 = sc.e1;
    CUDA_TRY(cudaMalloc(&buf, sizeof(double) * threads * (size_t)blocks));
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, 0));
        if (fp64)
            fma_peak_kernel<double><<<blocks, threads>>>(static_cast<double*>(buf), iters);
        else
            fma_peak_kernel<float><<<blocks, threads>>>(static_cast<float*>(buf), iters);
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    CUDA_TRY(cudaGetLastError());
    *tflops_out = best;
    return 0;
}

}  // extern "C"
