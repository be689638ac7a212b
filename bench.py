#!/usr/bin/env python
"""bench.py - MPC QP solves/sec (horizon 10) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (reference-class stand-in)

A "step" is one pass of the fused MPC tick (assembly + interior-point solve + torque map) over
one batch of synthetic randomised biped states; every rank owns an independent
shard (weak scaling, no data-path collective - SURVEY.md 8e); NCCL is used only for the timing
max and the final stats reduction.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import os

for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")  # the CPU legs run one process per core
if not os.environ.get("BENCH_KEEP_NCCL_DEBUG"):
    # NCCL prints its version banner on STDOUT at VERSION level and above (the level can also come from
    # /etc/nccl.conf); rank 0 must print exactly one JSON line, so debug output goes to stderr's file instead
    os.environ["NCCL_DEBUG"] = "WARN"

import argparse
import json
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpc_qp_solves_per_sec_h10"
UNIT = "solves/s"
PER_GPU_BATCH = 262144  # BASELINE.json configs[2]: 262,144 instances per GPU (weak scaling)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="instances per GPU per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch instances per GPU (default); strong: --batch instances in total, split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short h=30 and closed-loop rollout legs")
    return ap.parse_args()


def workload_config(batch, n_gpus, scaling="weak"):
    return {"workload": f"{batch} independent horizon-10 biped MPC ticks per GPU per step (BASELINE.json configs[2], "
                        f"{scaling} scaling; 85% walking / 15% standing, SURVEY.md 8d distribution)",
            "instances_per_gpu": batch, "horizon": 10, "parallelism": f"shard{n_gpus} (independent instances, no hot-path collective)",
            "l2_policy": "per-step inputs+outputs (~0.37 GB per GPU at 262144) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------
# algorithmic FLOPs (DESIGN.md "FLOP model"): 2 per FMA, mathematically required work only
# ------------------------------------------------------------------------------------------
def flops_per_solve(S, iters, LB=5, mb=11, h=10, stages=None, polish_rounds=1.15):
    n, m = LB * S, mb * S
    if stages is None:  # walking: one block per stage; standing: two per stage
        per_stage = max(1, round(S / h))
        stages = np.repeat(np.arange(h), per_stage)[:S]
    f_setup = 250.0 * h
    for jr in range(S):
        for jc in range(jr + 1):
            sr = stages[jr]
            f_setup += (h - 1 - sr) * (2 * 2 * 9 * LB + 2 * 3 * LB * LB) + 2 * 3 * LB * LB
    f_setup += S * h * 40.0
    # one Cholesky, Hc u, three solve pairs (predictor, corrector, Gondzio), row passes, block-diagonal update
    f_iter = n ** 3 / 3.0 + 8.0 * n * n + m * (8.0 * LB + 20.0) + S * mb * LB * (LB + 1.0)
    # polish round: one factorisation, two Hc products, one solve pair, null-space transform of every tile
    f_polish = n ** 3 / 3.0 + 8.0 * n * n + 4.0 * LB ** 3 * S * (S + 1) / 2.0
    return f_setup + iters * f_iter + (polish_rounds * f_polish if S > 0 else 0.0) + 600.0


def flops_per_solve_lane(S, iters, LB=5, mb=11, h=10, polish_rounds=1.15):
    """Restated for the stage-wise (Riccati) factorisation of the lane-per-robot kernel (SURVEY.md 8d asks for that when
    the factorisation differs): per stage with one stance foot the sweep is 1,211 FMA (P B 360, G + Cholesky
    125, F column operation 60, Y = inv(L) F 120, congruence 156, rank-5 update 390), a solve 224 FMA (backward + forward);
    one iteration = one sweep + two solves (no Gondzio corrector in this kernel) + the row passes + the stage weights."""
    m = mb * S
    per_foot = S / float(h)
    f_sweep, f_solve, f_grad = 2.0 * 1211.0 * h * per_foot, 2.0 * 224.0 * h * per_foot, 2.0 * 120.0 * h * per_foot
    f_setup = 250.0 * h + 40.0 * S * h + f_grad
    f_iter = f_sweep + 2.0 * f_solve + m * (8.0 * LB + 20.0) + S * mb * LB * (LB + 1.0)
    f_polish = f_sweep + f_solve + 3.0 * f_grad + 4.0 * LB ** 3 * S
    return f_setup + iters * f_iter + (polish_rounds * f_polish if S > 0 else 0.0) + 600.0


def lane_front_end_active(n, class_count, cls, h=10):
    """Mirror of the size gates in bmpc.cu / lane_tick_kernel: the lane-per-robot kernel takes a class only when the batch has
    >= 20,480 robots and the class >= 49,152 (walking) / 20,480 (standing) robots; BMPC_LANE_MIN overrides all three."""
    mode = os.environ.get("BMPC_LANE", "2")
    if h != 10 or mode == "0" or (cls == 1 and mode == "1"):
        return False
    ov = os.environ.get("BMPC_LANE_MIN")
    n_min, c_min = (int(ov), int(ov)) if ov else (20480, (49152, 20480)[cls])
    return n >= n_min and class_count >= c_min


def batch_flops(contact, iters, lane=False, lane_both=False):
    S = contact.reshape(contact.shape[0], -1).sum(axis=1).astype(int)
    per_stage = contact.reshape(contact.shape[0], -1, 2).sum(axis=2)
    total = {0: 0.0, 1: 0.0}
    h = contact.shape[1]
    for cls, feet in (((0, 1),) + (((1, 2),) if lane_both else ())) if lane else ():
        # robots with exactly `feet` stance feet in every stage are the lane kernel's (S = feet * h blocks = virtual stages)
        sel = (per_stage == feet).all(axis=1)
        base = flops_per_solve_lane(feet * h, 0.0, h=h)
        per_it = flops_per_solve_lane(feet * h, 1.0, h=h) - base
        total[cls] += sel.sum() * base + float(iters[sel].sum()) * per_it
        contact, iters, S, per_stage = contact[~sel], iters[~sel], S[~sel], per_stage[~sel]
    for s_val in np.unique(S):
        sel = S == s_val
        cls = 0 if s_val <= 10 else 1
        it_sum = float(iters[sel].sum())
        base = flops_per_solve(int(s_val), 0.0)
        per_it = flops_per_solve(int(s_val), 1.0) - base
        total[cls] += sel.sum() * base + it_sum * per_it
    return total


# ------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.dev, self.nv.NVML_CLOCK_SM)))
                bits = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU legs
# ------------------------------------------------------------------------------------------
def cpu_leg(per_core):
    from oracle import cpu_baseline
    r = cpu_baseline.run(per_core=per_core)
    return {"value": r["solves_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"{r['solves']} synthetic ticks ({per_core} per core, one process per core, 1 BLAS thread each): "
                      f"reference dense assembly + dense full-size interior point at cvxopt default tolerances "
                      f"(cvxopt-class stand-in; the real cvxopt is not installable) + lowLevelControl; "
                      f"{r['ms_per_solve']:.1f} ms per solve per core, of which assembly {r['ms_assembly']:.1f} ms",
            "per_core_value": r["per_core_solves_per_s"]}


def run_reference(args, rank):
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_leg(per_core=8)
        if i >= args.warmup:
            vals.append(r)
    value = float(np.mean([v["value"] for v in vals]))
    cores = vals[-1]["cores"]
    n_per_step = 8 * cores
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_per_step / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.batch, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# the other BASELINE.json configs, measured briefly after the headline (reported under "other_configs")
# ------------------------------------------------------------------------------------------
def other_configs(torch, dev, local_rank, rank, world, barrier, max_over_ranks, sum_over_ranks):
    from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
    tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    out = {}

    def timed(fn, reps):
        fn()  # warm-up (allocations, first launch)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / reps, r

    # configs[3]: horizon 30 (periodic gait extension for the walking instances; standing is reference-defined)
    n30 = 65536  # large enough for the lane-per-robot kernels (size gates 12,288 walking / 6,144 standing robots at h = 30)
    mpc30 = MPC(h=30)
    b = synth.make_batch(n30, shard_index=1000 + rank, mpc=mpc30, extend=True)
    s30 = BatchedMPC(mpc30, Biped(), max_batch=n30, device=local_rank, extend_gait=True)
    d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8),
         tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
    ms, r = timed(lambda: s30.step(*d), 2)
    st = r["status"].cpu().numpy()
    out["h30"] = {"workload": f"{n30} horizon-30 ticks per GPU (BASELINE.json configs[3]; 85% walking with the periodic gait "
                              "extension, 15% standing)", "solves_per_s": world * n30 / (ms * 1e-3), "ms_per_step": ms,
                  "mean_iters": float(r["iters"].float().mean().item()),
                  "not_optimal": int(sum_over_ranks(float((st != 0).sum())))}
    s30.close()

    # configs[4]: closed-loop rollout (MPC + J^T torques + SRB plant, rules R1-R7 of DESIGN.md 9)
    nr, ticks = 16384, 100
    sr = BatchedMPC(MPC(), synth.rollout_biped(), max_batch=nr, device=local_rank)
    rb = synth.make_rollout_batch(nr, shard_index=rank)
    for mode, warm in (("warm", True), ("cold", False)):
        def run():
            stt = [tn(rb["x"]), tn(rb["foot"]), tn(rb["tick"], torch.int32), tn(rb["gait"], torch.uint8), tn(rb["q"]), tn(rb["qd"])]
            return sr.rollout(*stt, ticks, warm_start=warm)
        ms, r = timed(run, 1)
        sn = BatchedMPC.rollout_stats(r["stats"])
        out["rollout_" + mode] = {"workload": f"{nr} robots x {ticks} ticks per GPU, closed loop (BASELINE.json configs[4] at a "
                                              f"tenth of its 1,000 ticks), warm_start={warm}; includes the H2D of the initial states",
                                  "robot_ticks_per_s": world * nr * ticks / (ms * 1e-3), "ms_per_tick": ms / ticks,
                                  "mean_iters": sn["mean_iters"], "warm_hit_rate": sn["warm_hit_rate"],
                                  "not_optimal": int(sum_over_ranks(float(sn["not_optimal"]))),
                                  "falls": int(sum_over_ranks(float(sn["falls"])))}
    sr.close()
    return out


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from biped_mpc_py_b200 import BatchedMPC, MPC, Biped, synth
    from tools.measure_peaks import measure

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL writes its version banner to STDOUT when the communicator is created; rank 0 must print exactly one
        # JSON line, so file descriptor 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    mpc, biped = MPC(), Biped()
    n = args.batch
    if args.scaling == "strong":  # fixed total work: contiguous shard of the same global batch
        from biped_mpc_py_b200.shard import shard_slice
        sl = shard_slice(args.batch, rank, world)
        n = sl.stop - sl.start
    batch = synth.make_batch(n, shard_index=rank, mpc=mpc, biped=biped)
    solver = BatchedMPC(mpc, biped, max_batch=n, device=local_rank)
    tn = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    d = dict(x_fb=tn(batch["x_fb"]), phase_k=tn(batch["phase_k"], torch.int32), t=tn(batch["t"]), foot=tn(batch["foot"]),
             contact=tn(batch["contact"], torch.uint8), q=tn(batch["q"]), qd=tn(batch["qd"]), pf_w=tn(batch["pf_w"]))

    def step():
        return solver.step(d["x_fb"], d["phase_k"], d["t"], d["foot"], d["contact"], d["q"], d["qd"], d["pf_w"])

    # ---- value: inputs resident in HBM, CUDA events on the launching stream --------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = solver.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    launches = solver.launch_count - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    ms_per_step = ms_total / args.steps
    total_n = int(sum_over_ranks(float(n)))
    value = total_n / (ms_per_step * 1e-3)

    iters = out["iters"].cpu().numpy()
    status = out["status"].cpu().numpy()
    resid = out["resid"].cpu().numpy()

    # ---- roofline of the two solve kernels (events inside the C ABI, same stream) ---------
    solver.enable_timing(True)
    kt = []
    for _ in range(max(3, min(args.steps, 10))):
        step()
        kt.append(solver.last_timing_ms())
    solver.enable_timing(False)
    kt = np.array(kt).mean(axis=0)  # classify, walking-class, standing-class
    cls_count = np.bincount((batch["contact"].reshape(n, -1).sum(axis=1) > 10).astype(int), minlength=2)
    lane = lane_front_end_active(n, int(cls_count[0]), 0)
    lane_both = lane and lane_front_end_active(n, int(cls_count[1]), 1)
    fl = batch_flops(batch["contact"], iters, lane=lane, lane_both=lane_both)
    fl_dense = batch_flops(batch["contact"], iters)
    peaks = measure(local_rank)
    kernels = []
    walk_name = ("lane_tick_kernel<10,1,5> (walking class: one THREAD per robot, stage-wise Riccati sweep; + collect + "
                 "mpc_tick2_kernel<10,10,5,32,8> for what it does not certify)") if lane else \
        "mpc_tick2_kernel<10,10,5,32,8> (<=10 stance foot-stages: walking class, one warp per robot, 8 robots per CTA)"
    stand_name = ("lane_tick_kernel<10,2,5> (standing class: one THREAD per robot, two virtual stages of one 5-input block per stage; "
                  "+ collect + mpc_tick2_kernel<10,20,5,128,1> for what it does not certify)") if lane_both else \
        "mpc_tick2_kernel<10,20,5,128,1> (11..20 stance foot-stages: standing class, one CTA per robot)"
    for cls, name in ((0, walk_name), (1, stand_name)):
        ach = fl[cls] / (kt[1 + cls] * 1e-3) / 1e12 if kt[1 + cls] > 0 else 0.0
        kernels.append({"kernel": name, "ms_per_launch": float(kt[1 + cls]), "algorithmic_gflop_per_launch": fl[cls] / 1e9,
                        "achieved_tflops": ach, "frac": ach / peaks["fp64_fma_tflops"],
                        # the same launch priced with the DENSE condensed-form count of the warp-per-robot kernels: the
                        # stage-wise form needs 1.7x (walking) / 5x (standing) fewer FLOPs for the same solves
                        "frac_at_dense_flop_count": (fl_dense[cls] / (kt[1 + cls] * 1e-3) / 1e12 / peaks["fp64_fma_tflops"]) if kt[1 + cls] > 0 else 0.0})
    dom = int(np.argmax(kt[1:]))
    traffic, secondary = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from one ncu capture of this batch size
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get("batch", -1)) == n:
            traffic = tj["kernels"][dom]["dram_bytes_per_launch"]
        kd = tj["kernels"][dom]
        if "smem_wavefronts_pct_of_peak" in kd:  # the nearest hardware limit per ncu (not measured live: profiler-only counters)
            secondary = {"what": "pipe utilisation of the dominant kernel from the committed ncu capture",
                         "smem_wavefronts_pct_of_peak": kd["smem_wavefronts_pct_of_peak"],
                         "fp64_pipe_pct_busy": kd["fp64_pipe_pct_busy"], "issue_slots_pct_busy": kd["issue_slots_pct_busy"],
                         "gcc_instruction_cache_busy_pct": kd.get("gcc_instruction_cache_busy_pct"),
                         "source": tj.get("pipe_source")}
    roofline = {"bound": "fp64_fma", "achieved": kernels[dom]["achieved_tflops"], "peak": peaks["fp64_fma_tflops"],
                "unit": "TFLOP/s", "frac": kernels[dom]["frac"], "traffic": traffic,
                "peak_source": "measured in this run by bmpc_measure_fma_peak (register-resident DFMA chains on all SMs); "
                               "MEASURED_PEAKS.json carries no FP64 CUDA-core figure",
                "fp32_fma_peak_tflops": peaks["fp32_fma_tflops"], "dominant_kernel": kernels[dom]["kernel"],
                "kernels": kernels, "classify_ms": float(kt[0]),
                "share_of_step": {"walking": float(kt[1] / kt.sum()), "standing": float(kt[2] / kt.sum())},
                "secondary": secondary}
    if traffic is not None and kernels[dom]["ms_per_launch"] > 0:
        # the lane-per-robot kernel streams its per-robot work arrays through DRAM: report that against the measured copy bandwidth
        hbm_peak = 6543.1
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        gbs = traffic / 1e9 / (kernels[dom]["ms_per_launch"] * 1e-3)
        roofline["hbm_view"] = {"what": "measured DRAM bytes of the dominant kernel (ncu, profiles/traffic.json) / its live duration: work-array "
                                        "streaming, not algorithmic bytes (~1.5 KB per robot)",
                                "dram_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak,
                                "dram_bytes_per_robot": traffic / max(1, int((batch["contact"].reshape(n, -1, 2).sum(axis=2) == 1).all(axis=1).sum()))
                                if lane else None}

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region --------
    tick = solver.pinned_tick(n, lowlevel=True, want_states=False)
    for k in ("x_fb", "foot", "q", "qd", "pf_w", "t", "phase_k", "contact"):
        tick.inputs[k][...] = batch[k]
    for _ in range(args.warmup):
        tick.run()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = tick.run()
        _ = float(res["tau"][0, 0])  # host read of the step's result
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": total_n * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(tick.h2d_bytes),
           "d2h_bytes_per_step": int(tick.d2h_bytes), "ms_per_step": 1e3 * e2e_s / args.steps,
           "api": "BatchedMPC.pinned_tick(n).run(): pinned host -> device, bmpc_step, device -> pinned host, sync"}

    # ---- stats reduction (the only collective): NCCL sum / max over ranks --------------------
    from biped_mpc_py_b200.shard import local_stats, reduce_stats
    stats = reduce_stats(*local_stats(status, iters, resid), device=dev)

    extra = None
    if not args.no_extra:
        extra = other_configs(torch, dev, local_rank, rank, world, barrier, max_over_ranks, sum_over_ranks)

    line = None
    if rank == 0:
        # single-instance latency through the reference-signature path (N=1, launch + copies + sync)
        latency = None
        if not args.no_latency:
            one = BatchedMPC(mpc, biped, max_batch=1, device=local_rank)
            t1 = one.pinned_tick(1)
            for k in ("x_fb", "foot", "q", "qd", "pf_w", "t", "phase_k", "contact"):
                t1.inputs[k][...] = batch[k][:1]
            lat = []
            for i in range(230):
                a = time.perf_counter()
                t1.run()
                lat.append(time.perf_counter() - a)
            lat = np.array(lat[30:]) * 1e3
            latency = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "samples": len(lat),
                       "what": "N=1 tick, pinned host in -> tau/controls on host, includes launch + copies + sync"}
            one.close()
            # the same call inside a caller-owned control loop for one walking robot (states replayed from a device rollout):
            # cold ticks vs ticks warm-started from the previous tick's certified active set (bmpc_warm_start)
            rb = synth.make_rollout_batch(64, shard_index=4)
            i = int(np.nonzero(rb["gait"] == 1)[0][0])
            loop = BatchedMPC(mpc, synth.rollout_biped(), max_batch=1, device=local_rank)
            nt = 160
            stt = [tn(rb[k][i:i + 1], dt) for k, dt in (("x", torch.float64), ("foot", torch.float64), ("tick", torch.int32),
                                                         ("gait", torch.uint8), ("q", torch.float64), ("qd", torch.float64))]
            lg = loop.rollout(*stt, nt, warm_start=False, n_log=1)
            torch.cuda.synchronize(dev)
            xl, fl = lg["x_log"].cpu().numpy()[:, 0], lg["foot_log"].cpu().numpy()[:, 0]
            tl = loop.pinned_tick(1)
            for warm in (False, True):
                loop.warm_start(warm)
                ll = []
                for k in range(nt):
                    T = int(rb["tick"][i]) + k
                    rows = (T + np.arange(10)) % 10
                    tl.inputs["x_fb"][0], tl.inputs["foot"][0], tl.inputs["pf_w"][0] = xl[k], fl[k], fl[k]
                    tl.inputs["q"][0], tl.inputs["qd"][0], tl.inputs["t"][0] = rb["q"][i], rb["qd"][i], T * mpc.dt
                    tl.inputs["phase_k"][0] = T % 10
                    tl.inputs["contact"][0] = np.stack([rows < 5, rows >= 5], axis=1).astype(np.uint8)
                    a = time.perf_counter()
                    tl.run()
                    ll.append(time.perf_counter() - a)
                latency["loop_warm_p50_ms" if warm else "loop_cold_p50_ms"] = float(np.percentile(np.array(ll[20:]) * 1e3, 50))
            latency["loop_what"] = ("one walking robot in a caller-owned control loop (160 ticks, states from a device rollout): "
                                    "cold ticks vs bmpc_warm_start ticks; same certified optimum")
            loop.close()
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_leg(per_core=12)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(n, world, args.scaling),
                "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "latency": latency, "other_configs": extra,
                "solver": {"mean_iters": float(stats["mean_iters"]), "max_iters": int(stats["iters_max"]),
                           "not_optimal": int(stats["not_optimal"]), "bad_input": int(stats["bad_input"]),
                           "max_mu": float(stats["mu_max"]), "max_rd": float(stats["rd_max"]),
                           "instances": int(stats["instances"])}}
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: spawn it ourselves
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
