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
# NCCL writes its debug output (communicator ranks, rings, NVLS) to STDOUT.  Rank 0 must print exactly one JSON line
# there, so the process's fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved stdout:
# NCCL_DEBUG=INFO stays visible (on stderr) for whoever wants to check the communicator.
# (Set, not defaulted: an inherited NCCL_DEBUG=WARN would hide the rank count the driver looks for.)
os.environ["NCCL_DEBUG"] = "INFO"
os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"

import argparse
import json
import sys
import threading
import time

sys.stdout.flush()
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

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
    """Restated for the stage-wise (Riccati) factorisation of the lane-per-robot kernels (SURVEY.md 8d asks for that when the
    factorisation differs).  Per virtual stage (one block of 5 inputs, structural input map):
    factor 1,176 FMA (P B 216, G 51, Cholesky 30, backward half of the solve 33 + 70, Yt = inv(L)(P B)' 180, rank-5 downdate
    of the cost-to-go 390, Y = Yt A 50, congruence A'PA 156); the other backward half 103 FMA, a forward half 103 FMA.
    One iteration = one factor sweep + three more half solves + the row work (priced with the dense count of the
    warp-per-robot model so that the two kernel families are comparable: the recomputation the lane kernel does instead of
    storing row quantities is NOT counted)."""
    m = mb * S
    f_factor, f_half, f_grad = 2.0 * 1176.0 * S, 2.0 * 103.0 * S, 2.0 * 120.0 * S
    f_setup = 250.0 * h + 40.0 * S * h + 2.0 * f_grad
    f_iter = f_factor + 3.0 * f_half + m * (8.0 * LB + 20.0) + S * mb * LB * (LB + 1.0)
    f_polish = f_factor + f_half + 2.0 * f_grad + 2.0 * (2.0 * 275.0 + 300.0) * S + 4.0 * LB ** 3 * S
    return f_setup + iters * f_iter + (polish_rounds * f_polish if S > 0 else 0.0) + 600.0


LANE_GATES = {10: (4096, 8192, 4096), 30: (1024, 2048, 1024)}  # batch, walking class, standing class (bmpc.cu::setup_lanes)


def lane_front_end_active(n, class_count, cls, h=10):
    """Mirror of the size gates in bmpc.cu / lane_tick_kernel: the lane-per-robot kernel takes a class only when the batch
    and the class are above the measured crossover against the warp-per-robot kernels."""
    n_min, c_min = LANE_GATES[h][0], LANE_GATES[h][1 + cls]
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
        # (the reported iteration count includes the last one, which only tests convergence and factors nothing)
        total[cls] += sel.sum() * base + float(np.maximum(iters[sel] - 1, 0).sum()) * per_it
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
            "per_core_value": r["per_core_solves_per_s"], "solves": r["solves"]}


REF_PER_CORE_PER_STEP = 128  # x (warmup + steps) >= 1,000 solves per process for the default --steps 10 --warmup 3


def run_reference(args, rank):
    """CPU arm: the oracle port (kind "port": reference-semantics dense assembly + dense interior point at cvxopt-default
    tolerances + lowLevelControl; cvxopt itself is not installable here) on every host core.  A step is a bounded sample of the
    workload: REF_PER_CORE_PER_STEP ticks per core."""
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_leg(per_core=REF_PER_CORE_PER_STEP)
        if i >= args.warmup:
            vals.append(r)
    value = float(np.mean([v["value"] for v in vals]))
    cores = vals[-1]["cores"]
    n_per_step = REF_PER_CORE_PER_STEP * cores
    cfg = workload_config(args.batch, args.gpus)
    cfg["workload"] = (f"CPU arm: {n_per_step} horizon-10 biped MPC ticks per step ({REF_PER_CORE_PER_STEP} per core on {cores} host cores) drawn "
                       f"from the same synthetic distribution as the GPU arm's {args.batch}-instance batch (BASELINE.json configs[2]; 85% walking / "
                       "15% standing, SURVEY.md 8d) - a bounded sample of that workload, not the whole batch")
    cfg["instances_per_step"] = n_per_step
    cfg["solves_per_process"] = REF_PER_CORE_PER_STEP * (args.warmup + args.steps)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_per_step / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


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
    n30 = 65536
    mpc30 = MPC(h=30)
    b = synth.make_batch(n30, shard_index=1000 + rank, mpc=mpc30, extend=True)
    s30 = BatchedMPC(mpc30, Biped(), max_batch=n30, device=local_rank, extend_gait=True)
    d = [tn(b["x_fb"]), tn(b["phase_k"], torch.int32), tn(b["t"]), tn(b["foot"]), tn(b["contact"], torch.uint8),
         tn(b["q"]), tn(b["qd"]), tn(b["pf_w"])]
    ms, r = timed(lambda: s30.step(*d), 5)
    st = r["status"].cpu().numpy()
    out["h30"] = {"workload": f"{n30} horizon-30 ticks per GPU (BASELINE.json configs[3]; 85% walking with the periodic gait "
                              "extension, 15% standing)", "solves_per_s": world * n30 / (ms * 1e-3), "ms_per_step": ms,
                  "mean_iters": float(r["iters"].float().mean().item()),
                  "not_optimal": int(sum_over_ranks(float((st != 0).sum())))}
    s30.close()

    # configs[4]: closed-loop rollout (MPC + J^T torques + SRB plant, rules R1-R7 of DESIGN.md 9)
    nr, ticks = 16384, 1000
    sr = BatchedMPC(MPC(), synth.rollout_biped(), max_batch=nr, device=local_rank)
    rb = synth.make_rollout_batch(nr, shard_index=rank)
    for mode, warm in (("warm", True), ("cold", False)):
        def short():
            stt = [tn(rb["x"]), tn(rb["foot"]), tn(rb["tick"], torch.int32), tn(rb["gait"], torch.uint8), tn(rb["q"]), tn(rb["qd"])]
            return sr.rollout(*stt, 10, warm_start=warm)
        short()  # warm-up: allocations, first launches
        stt = [tn(rb["x"]), tn(rb["foot"]), tn(rb["tick"], torch.int32), tn(rb["gait"], torch.uint8), tn(rb["q"]), tn(rb["qd"])]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = sr.rollout(*stt, ticks, warm_start=warm)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        sn = BatchedMPC.rollout_stats(r["stats"])
        out["rollout_" + mode] = {"workload": f"{nr} robots x {ticks} ticks per GPU, closed loop (BASELINE.json configs[4]), "
                                              f"warm_start={warm}; initial states resident on the device",
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
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)
        torch.cuda.synchronize(dev)

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

    # ---- roofline of the solve kernels (events inside the C ABI, same stream; with timing on, the kernels of a tick run
    #      one after the other so that the five intervals do not overlap) ----------------------------------------------
    solver.enable_timing(True)
    kt = []
    for _ in range(max(3, min(args.steps, 10))):
        step()
        kt.append(solver.last_timing_ms())
    solver.enable_timing(False)
    kt = np.array(kt).mean(axis=0)  # classify, lane walking, lane standing, warp-per-robot walking, warp-per-robot standing
    cls_count = np.bincount((batch["contact"].reshape(n, -1).sum(axis=1) > 10).astype(int), minlength=2)
    lane = lane_front_end_active(n, int(cls_count[0]), 0)
    lane_both = lane and lane_front_end_active(n, int(cls_count[1]), 1)
    fl = batch_flops(batch["contact"], iters, lane=lane, lane_both=lane_both)
    fl_dense = batch_flops(batch["contact"], iters)
    peaks = measure(local_rank)
    names = {("lane", 0): "lane_tick_kernel<10,1,RM> (walking class: one THREAD per robot, stage-wise Riccati sweeps, bmpc_lane.cuh)",
             ("lane", 1): "lane_tick_kernel<10,2,RM> (standing class: one THREAD per robot, two virtual stages of one 5-input block per stage)",
             ("warp", 0): "mpc_tick2_kernel<10,10,5,32,8> (walking class, one warp per robot, 8 robots per CTA; after a lane launch: collect + "
                          "what the lane kernel did not certify)",
             ("warp", 1): "mpc_tick2_kernel<10,20,5,128,1> (standing class, one CTA per robot; after a lane launch: collect + the rest)"}
    kernels = []
    for cls, is_lane in ((0, lane), (1, lane_both)):
        ms = float(kt[1 + cls]) if is_lane else float(kt[3 + cls])
        ach = fl[cls] / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        kernels.append({"kernel": names[("lane" if is_lane else "warp", cls)], "ms_per_launch": ms,
                        "algorithmic_gflop_per_launch": fl[cls] / 1e9, "achieved_tflops": ach, "frac": ach / peaks["fp64_fma_tflops"],
                        "fall_through_ms": float(kt[3 + cls]) if is_lane else None,
                        # the same launch priced with the DENSE condensed-form count of the warp-per-robot kernels: the
                        # stage-wise form needs 1.7x (walking) / 5x (standing) fewer FLOPs for the same solves
                        "frac_at_dense_flop_count": (fl_dense[cls] / (ms * 1e-3) / 1e12 / peaks["fp64_fma_tflops"]) if ms > 0 else 0.0})
    dom = int(np.argmax([k["ms_per_launch"] for k in kernels]))
    traffic, secondary = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from one ncu capture of this batch size
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        tk = tj.get("kernels", {})
        kd = tk.get(("lane_" if (lane, lane_both)[dom] else "warp_") + ("walking", "standing")[dom]) if isinstance(tk, dict) else None
        if kd:
            if int(tj.get("batch", -1)) == n:
                traffic = kd["dram_bytes_per_launch"]
            secondary = {"what": "pipe utilisation of the dominant kernel from the committed ncu capture (profiler-only counters, not measured live)",
                         "fp64_pipe_pct_busy": kd.get("fp64_pipe_pct_busy"), "issue_slots_pct_busy": kd.get("issue_slots_pct_busy"),
                         "threads_per_instruction": kd.get("threads_per_instruction"),
                         "l2_hit_rate_pct": kd.get("l2_hit_rate_pct"), "source": tj.get("source")}
    ksum = float(kt.sum())
    roofline = {"bound": "fp64_fma", "achieved": kernels[dom]["achieved_tflops"], "peak": peaks["fp64_fma_tflops"],
                "unit": "TFLOP/s", "frac": kernels[dom]["frac"], "traffic": traffic,
                "peak_source": "measured in this run by bmpc_measure_fma_peak (register-resident DFMA chains on all SMs); "
                               "MEASURED_PEAKS.json carries no FP64 CUDA-core figure",
                "fp32_fma_peak_tflops": peaks["fp32_fma_tflops"], "dominant_kernel": kernels[dom]["kernel"],
                "kernels": kernels, "classify_ms": float(kt[0]),
                "serial_ms": {"classify": float(kt[0]), "lane_walking": float(kt[1]), "lane_standing": float(kt[2]),
                              "warp_walking": float(kt[3]), "warp_standing": float(kt[4]), "sum": ksum,
                              "note": "per-kernel durations with the kernels of a tick run one after the other; in the timed region the two "
                                      "class chains run concurrently on two streams, so ms_per_step is below this sum"},
                "share_of_step": {"lane_walking": float(kt[1] / ksum), "lane_standing": float(kt[2] / ksum),
                                  "warp_walking": float(kt[3] / ksum), "warp_standing": float(kt[4] / ksum)},
                "whole_step_frac": float(sum(fl.values()) / (ms_per_step * 1e-3) / 1e12 / peaks["fp64_fma_tflops"]),
                "secondary": secondary}
    if traffic is not None and kernels[dom]["ms_per_launch"] > 0:
        # the lane-per-robot kernels stream their per-robot block records through L2 / DRAM: report that against the measured copy bandwidth
        hbm_peak = 6543.1
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        gbs = traffic / 1e9 / (kernels[dom]["ms_per_launch"] * 1e-3)
        roofline["hbm_view"] = {"what": "measured DRAM bytes of the dominant kernel (ncu, profiles/traffic.json) / its live duration: block-record "
                                        "streaming, not algorithmic bytes (~1.5 KB per robot)",
                                "dram_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak,
                                "dram_bytes_per_robot": traffic / max(1, int(cls_count[dom]))}

    # ---- e2e: host buffers through the public API, H2D + D2H of every step inside the timed region -----------------
    # Two ticks (slots) of two chunks each: step i+1 is launched before the host reads the result of step i, so the copies of
    # one step run under the kernels of the other (BatchedMPC.chunked_tick, api.py).  Every step copies its packed inputs from
    # pinned host memory and its packed results back to pinned host memory, one cudaMemcpyAsync per chunk and direction.
    keys = ("x_fb", "foot", "q", "qd", "pf_w", "t", "phase_k", "contact")
    n_chunks = 2 if n >= 2 * 16384 else 1
    ticks = [solver.chunked_tick(n, n_chunks, lowlevel=True, want_states=False, slot=k) for k in range(2)]
    for tk in ticks:
        tk.set_inputs(**{k: batch[k] for k in keys})
    for _ in range(max(1, args.warmup)):
        for tk in ticks:
            tk.run()
    barrier()
    t0 = time.perf_counter()
    ticks[0].launch()
    for i in range(1, args.steps):
        ticks[i % 2].launch()
        res = ticks[(i - 1) % 2].wait()
        _ = float(res[0]["tau"][0, 0])  # host read of the step's result
    res = ticks[(args.steps - 1) % 2].wait()
    _ = float(res[0]["tau"][0, 0])
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    # the same through ONE synchronous call per step (no overlap between steps)
    t0 = time.perf_counter()
    for _ in range(min(args.steps, 5)):
        res = ticks[0].run()
        _ = float(res[0]["tau"][0, 0])
    sync_s = max_over_ranks(time.perf_counter() - t0) / min(args.steps, 5)
    e2e = {"value": total_n * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(ticks[0].h2d_bytes),
           "d2h_bytes_per_step": int(ticks[0].d2h_bytes), "ms_per_step": 1e3 * e2e_s / args.steps,
           "api": f"BatchedMPC.chunked_tick(n, chunks={n_chunks}, slot=0|1): launch() of step i+1 before wait() of step i; per chunk one "
                  "cudaMemcpyAsync pinned host -> device, bmpc_step, one cudaMemcpyAsync device -> pinned host",
           "synchronous_single_call": {"value": total_n / sync_s, "ms_per_step": 1e3 * sync_s,
                                       "api": f"chunked_tick(n, chunks={n_chunks}).run(): copies in, kernels, copies out, wait - one step at a time"}}

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
            cpu = cpu_leg(per_core=1000)  # >= 1,000 solves per process: ~20 s on every host core
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(n, world, args.scaling),
                "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "latency": latency, "other_configs": extra,
                "solver": {"mean_iters": float(stats["mean_iters"]), "max_iters": int(stats["iters_max"]),
                           "not_optimal": int(stats["not_optimal"]), "bad_input": int(stats["bad_input"]),
                           "max_mu": float(stats["mu_max"]), "max_rd": float(stats["rd_max"]),
                           "instances": int(stats["instances"])}}
        emit(line)
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
