"""Refresh profiles/ from a gpurun_out/ measurement set taken with the lane-per-robot front end on.
usage: python tools/update_profiles_lane.py <launch tag, e.g. v13> <traffic/ncu tag, e.g. 12>
(expects gpurun_out/launches_<tag>.csv, traffic<N>_262144.csv, prof_r1_v<N>.ncu-rep)"""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
ltag, num = sys.argv[1], sys.argv[2]
shutil.copy(os.path.join(G, f"launches_{ltag}.csv"), os.path.join(P, f"r1_launches_bench_{ltag}.csv"))
shutil.copy(os.path.join(G, f"traffic{num}_262144.csv"), os.path.join(P, "r1_traffic_262144.csv"))
met = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"prof_r1_v{num}.ncu-rep")],
                     capture_output=True, text=True).stdout
open(os.path.join(P, f"r1_ncu_v{num}_metrics.txt"), "w").write(met)
pipes = {}
for blk in met.split("-----")[1:]:
    name = re.search(r"Kernel Name\s+(.*)", blk).group(1).strip()
    g = lambda key: float(re.search(re.escape(key) + r" \S* ?([0-9.]+)", blk).group(1))
    pipes[name.split("(")[0].strip()] = dict(
        smem_wavefronts_pct_of_peak=round(g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), 1),
        fp64_pipe_pct_busy=round(g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 1),
        issue_slots_pct_busy=round(g("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1))
rows = [r for r in csv.reader(open(os.path.join(P, "r1_traffic_262144.csv"))) if len(r) > 10 and r[0].isdigit()]
k = collections.OrderedDict()
for r in rows:
    k.setdefault(r[4], {})[r[12]] = float(r[14].replace(",", ""))
# bench.py reads kernels[0] for the walking class and kernels[1] for the standing class
order = [n for n in k if "lane_tick_kernel<10, 1" in n] + [n for n in k if "lane_tick_kernel<10, 2" in n] + [n for n in k if "mpc_tick2" in n]
out = {"batch": 262144, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gcc__* --clock-control none "
       f"-k regex:lane_tick|mpc_tick2 -c 4 python tools/prof_driver.py 262144 1 (profiles/r1_traffic_262144.csv, v{num})",
       "pipe_source": f"ncu --set full, 65,536-robot batch, profiles/r1_ncu_v{num}_metrics.txt", "kernels": []}
for name in order:
    m = k[name]
    d = {"kernel": name, "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
         "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"], "ncu_duration_ms": m["gpu__time_duration.sum"] / 1e6,
         "gcc_instruction_cache_busy_pct": round(100 * m["gcc__cycles_active.avg"] / m["gcc__cycles_elapsed.avg"], 1),
         "gcc_requests": m["gcc__cache_requests.sum"]}
    d.update(pipes.get(name.split("(")[0].strip(), {}))
    out["kernels"].append(d)
json.dump(out, open(os.path.join(P, "traffic.json"), "w"), indent=1)
for d in out["kernels"]:
    print({a: (round(b, 2) if isinstance(b, float) else b) for a, b in d.items()})
rows = [r for r in csv.reader(open(os.path.join(P, f"r1_launches_bench_{ltag}.csv"))) if len(r) > 10 and r[0].isdigit()]
t = collections.Counter()
for r in rows:
    t[r[4][:60]] += float(r[-1])
tot = sum(t.values())
for n, v in t.most_common(8):
    print(f"{100 * v / tot:6.2f}%  {n}")
