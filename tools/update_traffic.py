"""profiles/traffic.json from one `ncu --set full` capture of the lane-per-robot kernels at the bench batch size: DRAM bytes per
launch (bench.py's roofline.traffic) and the pipe utilisation figures bench.py reports as `roofline.secondary`.
usage: python tools/update_traffic.py <report.ncu-rep> <batch> "<how it was captured>" """
import csv
import io
import json
import os
import subprocess
import sys

rep, batch, how = sys.argv[1], int(sys.argv[2]), sys.argv[3]
# (a .csv argument: the raw page already exported on the GPU box - full reports of six launches exceed what comes back from there)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
f = lambda r, k: float(r[col[k]].replace(",", "")) if k in col and r[col[k]] not in ("", "n/a") else None
out = {"batch": batch, "source": how, "kernels": {}}
for r in data:
    name = r[col["Kernel Name"]]
    if "lane_tick_kernel" in name:
        key = "lane_walking" if "<10, 1," in name or "<30, 1," in name else "lane_standing"
    elif "mpc_tick2_kernel" in name:
        key = "warp_walking" if ", 32, 8" in name or "10, 10," in name else "warp_standing"
    else:
        continue
    if key in out["kernels"]:
        # the lane kernels of a class are launched three times per tick (first pass, interior-point pass, polish pass over the
        # parked robots): the later launches add to the class's traffic and duration; the pipe figures stay the first pass's
        if key.startswith("lane_"):
            unit = lambda k: rows[1][col[k]]
            scale = lambda k: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[unit(k)]
            kd = out["kernels"][key]
            if kd.get("passes", 1) < 3:
                rd, wr = f(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum"), f(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
                ms = f(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit("gpu__time_duration.sum")]
                kd["dram_bytes_read"] += rd
                kd["dram_bytes_write"] += wr
                kd["dram_bytes_per_launch"] += rd + wr
                kd["ncu_duration_ms"] += ms
                kd.setdefault("pass_ms", [kd["ncu_duration_ms"] - ms]).append(ms)
                kd.setdefault("pass_dram_bytes", [kd["dram_bytes_per_launch"] - rd - wr]).append(rd + wr)
                kd["passes"] = kd.get("passes", 1) + 1
        continue
    unit = lambda k: rows[1][col[k]]
    scale = lambda k: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[unit(k)]
    rd, wr = f(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum"), f(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    hit, miss = f(r, "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"), f(r, "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum")
    out["kernels"][key] = {
        "kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
        "ncu_duration_ms": f(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit("gpu__time_duration.sum")],
        "fp64_pipe_pct_busy": f(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_slots_pct_busy": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "threads_per_instruction": f(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "registers_per_thread": f(r, "launch__registers_per_thread"),
        "l2_hit_rate_pct": 100.0 * hit / (hit + miss) if hit is not None and miss else None,
        "local_load_sectors": f(r, "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
        "local_store_sectors": f(r, "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
        "gcc_instruction_requests_pct_of_peak": f(r, "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"),
        "warp_instructions": f(r, "smsp__inst_executed.sum"),
    }
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
