/*
 * biped_mpc_b200 - C ABI of the B200-native batched HECTOR-style force-and-moment MPC.
 *
 * Drop-in boundary for the hot path of /root/reference/bipedalLocomotionMPC.py ("MPC.py"):
 *   solve_mpc(x_fb, t, foot, mpc, biped, contact) -> (states, controls)      MPC.py:187-304
 *   lowLevelControl(x_fb, t, pf_w, q, qd, mpc, biped, contact, u) -> tau     MPC.py:444-470
 * The reference has no FFI layer (two plain Python functions called at MPC.py:487 and
 * MPC.py:494); these entry points are what a ctypes binding for that path binds, batched
 * over N independent robots.  Plain pointers and sizes only - no torch / C++ types.
 *
 * Conventions
 *   - All real arrays are float64, row-major, contiguous, DEVICE pointers owned by the
 *     caller (torch tensor.data_ptr()), unless the function name ends in _host.
 *   - Every call returns 0 on success, non-zero on failure; bmpc_last_error() gives the text.
 *   - Calls are asynchronous on the given CUDA stream (a cudaStream_t passed as void*;
 *     NULL = legacy default stream).  One handle per host thread / stream.
 *   - State x = [euler(3), pos(3), omega(3), vel(3)] (MPC.py:9), input
 *     u = [f1(3), f2(3), m1(3), m2(3)] (MPC.py:10).
 */
#ifndef BIPED_MPC_B200_H
#define BIPED_MPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMPC_ABI_VERSION 2

/* Mirror of `class MPC` (MPC.py:22-32) + `class Biped` (MPC.py:34-48) + solver options. */
typedef struct bmpc_params {
    int32_t h;             /* horizon: 10 or 30                                  MPC.py:24 */
    int32_t extend_gait;   /* 1: periodic foot-reference extension for h != 10 (new
                              behaviour, the reference raises IndexError there)           */
    double dt;             /* MPC.py:25 */
    double x_cmd[12];      /* MPC.py:26 */
    double Q[13];          /* MPC.py:27 */
    double R[12];          /* MPC.py:28 */
    double kv;             /* MPC.py:29 */
    double kp[9];          /* row-major 3x3, MPC.py:30 */
    double kd[9];          /* row-major 3x3, MPC.py:31 */
    double swing_height;   /* MPC.py:32 */
    double mass;           /* MPC.py:36 */
    double inertia[9];     /* row-major 3x3, MPC.py:37-39 */
    double lt, lh;         /* MPC.py:40-41 (the 0.01 / 0.02 margins of MPC.py:254-255 are applied inside) */
    double g;              /* MPC.py:42 */
    double hip_offset[3];  /* MPC.py:43 (used by bmpc_foot_positions only) */
    double mu;             /* MPC.py:44 */
    double f_max[3], f_min[3], tau_max[3], tau_min[3]; /* MPC.py:45-48 */
    /* solver options (0 selects the default in brackets) */
    int32_t max_iter;      /* [40] interior-point iteration cap                                    */
    int32_t polish;        /* reserved (the active-set polish is always on)                        */
    double mu_tol;         /* [1e-7] complementarity mu <= mu_tol*(1+|g|inf) hands over to the polish */
    double rd_tol;         /* [10]   stationarity residual <= rd_tol * that mu target               */
} bmpc_params;

typedef struct bmpc_handle bmpc_handle;

/* status[] values written per instance */
#define BMPC_STATUS_OPTIMAL   0   /* exact optimum: active-set polish certified (KKT)    */
#define BMPC_STATUS_MAXITER   1   /* iteration cap hit; best iterate returned            */
#define BMPC_STATUS_NUMERIC   2   /* factorisation broke down and the polish did not certify; iterate returned */
#define BMPC_STATUS_BADINPUT  3   /* NaN/Inf in inputs or pitch = +-pi/2 (MPC.py:160-164 singular) */

/* Create a solver bound to `device` for batches up to `max_batch`.  Replaces the
 * construction of `mpc = MPC(); biped = Biped()` (MPC.py:475-476). */
int bmpc_create(const bmpc_params* params, int device, int max_batch, bmpc_handle** out);
int bmpc_destroy(bmpc_handle* h);

/* One MPC tick for N robots = solve_mpc (MPC.py:487) + lowLevelControl on controls[0]
 * (MPC.py:493-494), fused.
 *   in : x_fb[N,12]  phase_k[N] (int32: int(t // dt) % h computed on the host with the
 *        reference's float expression, MPC.py:56,99)  t_swing[N] (the t of MPC.py:436)
 *        foot[N,6] (MPC.py:479)  contact[N,h,2] uint8 (MPC.py:482-484)  q[N,10]  qd[N,10]
 *        pf_w[N,6]
 *   out: controls[N,h,12]  states[N,h,13] (nullable)  tau[N,10]  status[N]  iters[N]
 *        fric_active[N,h] (nullable; bit 4*leg+r = friction row r of MPC.py:220-229 active)
 *        resid[N,2] (nullable; complementarity mu - 0 once polished - and the last stationarity residual) */
int bmpc_step(bmpc_handle* h, int n,
              const double* x_fb, const int32_t* phase_k, const double* t_swing,
              const double* foot, const uint8_t* contact,
              const double* q, const double* qd, const double* pf_w,
              double* controls, double* states, double* tau,
              int32_t* status, int32_t* iters, uint8_t* fric_active, double* resid,
              void* stream);

/* solve_mpc only (MPC.py:187-304). */
int bmpc_solve(bmpc_handle* h, int n,
               const double* x_fb, const int32_t* phase_k, const double* foot, const uint8_t* contact,
               double* controls, double* states,
               int32_t* status, int32_t* iters, uint8_t* fric_active, double* resid,
               void* stream);

/* lowLevelControl only (MPC.py:444-470): u0[N,12] -> tau[N,10].  contact0[N,2] = contact[0,:]. */
int bmpc_lowlevel(bmpc_handle* h, int n,
                  const double* x_fb, const double* t_swing, const double* pf_w,
                  const double* q, const double* qd, const uint8_t* contact0, const double* u0,
                  double* tau, void* stream);

/* getFootPositionWorld (MPC.py:406-424) for N robots: x_fb[N,12], q[N,10] -> pf_w[N,6]. */
int bmpc_foot_positions(bmpc_handle* h, int n, const double* x_fb, const double* q, double* pf_w, void* stream);

/* Closed-loop batched rollout (BASELINE.json configs[4]): `ticks` control ticks for N
 * robots, each tick = bmpc_step + one step of the reference's own discretised single-rigid-body
 * model x+ = A_0 [x;1] + B_0 u_0 (MPC.py:148-185, 206-208) + gait clock + touchdown foothold.
 * The reference has no loop (MPC.py:475-495 runs one tick); the rules R1-R6 are stated in
 * DESIGN.md section 9 and restated for the CPU in oracle/rollout.py.
 *   in/out: x[N,12]  foot[N,6] (= pf_w)  tick[N] (int32 gait clock; phase = tick % 10, MPC.py:56-58)
 *   in    : gait[N] uint8 (1 walking / 0 standing, MPC.py:18)  q[N,10]  qd[N,10] (held fixed)
 *           warm_start != 0: after the first tick, each solve starts from the previous tick's
 *           certified active set shifted by one stage and goes straight to the active-set
 *           polish (falls back to the cold interior point when that does not certify; results are
 *           the same certified optimum either way - the reference solves cold, MPC.py:297)
 *   logs  : first n_log robots: x_log[ticks+1,n_log,12] foot_log[ticks+1,n_log,6]
 *           u0_log[ticks,n_log,12] tau_log[ticks,n_log,10] (all nullable when n_log == 0)
 *   stats : uint64[8] DEVICE, accumulated (caller zeroes): [0] sum of interior-point iterations
 *           [1] ticks not certified optimal [2] bad-input ticks [3] max iterations [4] robot-ticks
 *           [5] ticks solved by the warm polish alone (0 iterations) [6] falls (rule R7).  Nullable. */
int bmpc_rollout(bmpc_handle* h, int n, int ticks,
                 double* x, double* foot, int32_t* tick, const uint8_t* gait,
                 const double* q, const double* qd, int warm_start,
                 int n_log, double* x_log, double* foot_log, double* u0_log, double* tau_log,
                 uint64_t* stats, void* stream);

/* Warm start across calls for a caller-owned control loop (the reference solves cold every
 * call, MPC.py:297).  mode 1: every following bmpc_step / bmpc_solve assumes that robot i of the batch is the same
 * robot ONE control tick later (horizon shifted by one stage) and starts from the active set certified by the
 * previous call (first call: cold).  A guess that does not certify falls back to the cold solve, so results are the
 * same certified optimum either way.  mode 0: off (default).  mode 2: forget the stored sets (after a reset / a
 * jump in time), stay on. */
int bmpc_warm_start(bmpc_handle* h, int mode);

/* Debug / parity: the contact-reduced condensed QP of instance `index` of the last
 * bmpc_step/bmpc_solve inputs.  Hc_out[nmax*nmax] row-major (nmax = 12*h), g_out[nmax],
 * n_out = number of reduced variables, all DEVICE pointers.  Synchronous. */
int bmpc_debug_assemble(bmpc_handle* h,
                        const double* x_fb, const int32_t* phase_k, const double* foot, const uint8_t* contact,
                        double* Hc_out, double* g_out, int32_t* n_out, void* stream);

/* Tunables of the kernel dispatch (integers; the defaults need no call).  They replace the environment
 * variables of earlier builds; results are the same certified optimum under every setting.
 *   "lane_mode"         2 (default) the lane-per-robot kernels take both instance classes of throughput
 *                       batches, 1 the walking class only, 0 warp-per-robot kernels only
 *   "lane_min"          overrides the three size gates of the lane-per-robot kernels (batch and class sizes
 *                       below which a class stays on the warp-per-robot kernel); -1 restores the defaults
 *   "lane_ctas_per_sm"  resident CTAs per SM of the lane-per-robot kernels (0 / -1: as many as fit)
 *   "lane_warps"        warps (32 robots each) per CTA of the lane-per-robot kernels (default 4)
 *   "lane_ctas_standing" resident CTAs per SM of the standing class's lane kernel only (measurement: co-residency of the classes)
 *   "lane_prefetch"     0 disables the bulk L2 prefetch of the lane-per-robot kernels (measurement only)
 *   "lane_sync"         lockstep of the warps of a lane-kernel CTA: 2 (default) barrier per stage of every sweep,
 *                       1 per iteration, 0 independent warps; + 4: the polish of every warp on its own; + 8: one barrier
 *                       per sweep instead of per stage (measurement only: all measured slower than the default)
 *   "lane_ipm_inline"   interior-point iterations of the lane kernels' first pass before a robot that has not converged
 *                       is parked for the interior-point pass (default 9 at h = 10, 0 = never at h = 30)
 *   "lane_inline_rounds" polish rounds of the first pass before a robot that needs more is parked for the polish pass
 *                       (default 1; 0: no later passes at all)
 *   "lane_defer_min"    smallest class for which robots are parked (default: two waves of resident slices; -1 restores it)
 *   "lane_defer_cap"    size of the two park stores in robots (default: a quarter of max_batch, at least 4,096); a robot that
 *                       finds its store full carries on in the pass it is in
 *   "polish_rounds"     budget of polish rounds per attempt (default 4 at h = 10, 16 at h = 30)
 *   "lowlat"            0 disables the 128-thread low-latency kernel used for batches <= 8
 * With "lane_min" = 1, "lane_defer_min" = 1 and "lowlat" = 0 every robot takes the same code path whatever batch it
 * arrives in: results are bit-identical for any sharding of the robots over calls or GPUs ("lane_mode" = 0 with
 * "lowlat" = 0 does the same with the warp-per-robot kernels); with the default size gates the kernel family, and with it
 * the last bits of a result, depend on the batch.
 * Synchronises the device (the lane workspace is re-sized). */
int bmpc_set_option(bmpc_handle* h, const char* name, int value);

/* Number of kernels this handle has launched so far (for bench.py's gpu_launches). */
int64_t bmpc_launch_count(const bmpc_handle* h);

/* Per-kernel device timing of the last bmpc_step/bmpc_solve (CUDA events on the launching
 * stream): ms5 = {classify, lane-per-robot kernels of the walking class (first pass + interior-point pass + polish
 * pass), lane-per-robot kernels of the standing class (each 0 when the batch is below the size gates), walking-class warp-per-robot kernel
 * (<= h stance foot-stages; after a lane launch: the collect step + what that did not certify),
 * standing-class warp-per-robot kernel (+ the h = 30 dense re-solve)}.  While timing is enabled the
 * kernels of a tick run one after the other (normally the two classes run concurrently on two
 * streams), so the five intervals do not overlap.  bmpc_last_timing blocks until the tick has finished. */
int bmpc_enable_timing(bmpc_handle* h, int enable);
int bmpc_last_timing(bmpc_handle* h, float* ms5);

/* Measured CUDA-core FMA peak of the device (roofline denominator): fp64 != 0 selects
 * double precision.  Runs a register-resident FMA chain on every SM; result in TFLOP/s. */
int bmpc_measure_fma_peak(int device, int fp64, double* tflops_out);

const char* bmpc_last_error(void);
int bmpc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BIPED_MPC_B200_H */
