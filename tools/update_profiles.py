"""Refresh profiles/ from a gpurun_out/ measurement set.
usage: python tools/update_profiles.py <tag, e.g. v10>   (expects gpurun_out/launches_<tag>.csv, traffic<N>_262144.csv, prof_r1_<tag>.ncu-rep)"""
import csv, glob, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
num = re.sub(r"\D", "", tag)
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for old in glob.glob(os.path.join(P, "r1_launches_bench_v*.csv")) + glob.glob(os.path.join(P, "r1_ncu_v[0-9]*_metrics.txt")):
    if re.search(r"_v(4|6)[._]", old):
        continue  # keep the historical captures the summary cites
    os.remove(old)
shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"r1_launches_bench_{tag}.csv"))
shutil.copy(os.path.join(G, f"traffic{num}_262144.csv"), os.path.join(P, "r1_traffic_262144.csv"))
met = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"prof_r1_{tag}.ncu-rep")],
                     capture_output=True, text=True).stdout
open(os.path.join(P, f"r1_ncu_{tag}_metrics.txt"), "w").write(met)
pipes = []
for blk in met.split("-----")[1:]:
    g = lambda k: float(re.search(re.escape(k) + r" \S* ?([0-9.]+)", blk).group(1))
    pipes.append(dict(smem_wavefronts_pct_of_peak=round(g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), 1),
                      fp64_pipe_pct_busy=round(g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 1),
                      issue_slots_pct_busy=round(g("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1)))
rows = [r for r in csv.reader(open(os.path.join(P, "r1_traffic_262144.csv"))) if len(r) > 10 and r[0].isdigit()]
k = {}
for r in rows:
    k.setdefault(r[4], {})[r[12]] = float(r[14].replace(",", ""))
out = {"batch": 262144, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gcc__* --clock-control none "
       f"-k regex:mpc_tick2 -c 2 python tools/prof_driver.py 262144 1 (profiles/r1_traffic_262144.csv, {tag})",
       "pipe_source": f"ncu --set full, 8,192-robot batch, profiles/r1_ncu_{tag}_metrics.txt", "kernels": []}
for i, (name, m) in enumerate(k.items()):
    d = {"kernel": name, "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
         "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"], "ncu_duration_ms": m["gpu__time_duration.sum"] / 1e6,
         "gcc_instruction_cache_busy_pct": round(100 * m["gcc__cycles_active.avg"] / m["gcc__cycles_elapsed.avg"], 1),
         "gcc_requests": m["gcc__cache_requests.sum"]}
    d.update(pipes[i] if i < len(pipes) else {})
    out["kernels"].append(d)
json.dump(out, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(out["kernels"], indent=1))
# launch shares
rows = [r for r in csv.reader(open(os.path.join(P, f"r1_launches_bench_{tag}.csv"))) if len(r) > 10 and r[0].isdigit()]
t = {}
for r in rows:
    t[r[4][:48]] = t.get(r[4][:48], 0.0) + float(r[14].replace(",", ""))
tot = sum(t.values())
for n, v in sorted(t.items(), key=lambda kv: -kv[1]):
    print(f"{100 * v / tot:6.2f}%  {n}")
