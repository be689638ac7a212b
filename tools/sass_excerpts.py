"""Write profiles/<round>_sass.md: per kernel, instruction counts by class and the hot inner loops (trailing tile update of the
block Cholesky; the TMA bulk loads and mbarrier waits), taken from `cuobjdump -sass` of the in-tree library.
usage: python tools/sass_excerpts.py > profiles/r1_sass.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "biped_mpc_py_b200", "csrc", "libbiped_mpc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines()
funcs, cur = collections.OrderedDict(), None
for ln in sass:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("# SASS of the sm_100a kernels (cuobjdump -sass of libbiped_mpc_b200.so)\n")
print("Instruction classes per kernel (static counts; helper device functions are part of the kernel's text):\n")
print("| kernel | instructions | code KB | DFMA | DMUL/DADD | LDS | STS | LDG/LD | STG/ST | SHFL | UBLKCP (TMA bulk) | SYNCS (mbarrier) | BAR | MUFU |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, ins in funcs.items():
    c = collections.Counter()
    for _, t in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
        base = op.split(".")[0]
        c[base] += 1
    d = demangle(name)
    d = re.sub(r"\(.*", "", d).replace("void bmpc::", "").replace("bmpc::", "")
    print(f"| `{d}` | {len(ins)} | {len(ins) * 16 / 1024:.0f} | {c['DFMA']} | {c['DMUL'] + c['DADD']} | {c['LDS']} | {c['STS']} | "
          f"{c['LDG'] + c['LD']} | {c['STG'] + c['ST']} | {c['SHFL']} | {c['UBLKCP']} | {c['SYNCS']} | {c['BAR']} | {c['MUFU']} |")
# excerpts from the shipped walking-class kernel
key = ([n for n in funcs if "Li10ELi10ELi5ELi32ELi8ELb0" in n] + [n for n in funcs if "Li10ELi10ELi5ELi32E" in n])[0]
ins = funcs[key]
print(f"\n## Excerpts from `{demangle(key).split('(')[0]}` (walking class)\n")
idx = [i for i, (_, t) in enumerate(ins) if "UBLKCP" in t]
print("Input staging: TMA 1-D bulk copies (`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`) of x_fb / foot / q / qd / pf_w,")
print("completion on an mbarrier (`SYNCS.ARRIVE.TRANS64` = `mbarrier.arrive.expect_tx`, `SYNCS.PHASECHK` = `mbarrier.try_wait.parity`):\n\n```")
for a, t in ins[max(0, idx[0] - 14): idx[0] + 3]:
    print(f"/*{a:05x}*/ {t}")
print("```\n")
# densest DFMA window = the rolled trailing update of tile_factor (25 DFMA per k)
best, bi = -1, 0
W = 48
isf = [1 if "DFMA" in t else 0 for _, t in ins]
s = sum(isf[:W])
for i in range(len(ins) - W):
    if s > best:
        best, bi = s, i
    s += isf[i + W] - isf[i]
print(f"Densest FP64 window ({best} DFMA in {W} instructions): the rolled k-loop of the trailing tile update `A(jr,jc2) -= L(jr,jc) L(jc2,jc)'`")
print("(`tile_factor`, bmpc_tick.cuh): 10 `LDS.64` of one column of each panel tile feed 25 register-accumulator DFMAs:\n\n```")
for a, t in ins[bi: bi + W]:
    print(f"/*{a:05x}*/ {t}")
print("```")

# lane-per-robot kernel: densest DFMA window (the rank-5 update of the packed cost-to-go in the register-blocked sweep) and the
# streaming accesses of the stored factor
lk = [n for n in funcs if "lane_tick_kernelILi10ELi1ELi5" in n]
if lk:
    ins = funcs[lk[0]]
    print(f"\n## Excerpts from `{demangle(lk[0]).split('(')[0]}` (lane-per-robot kernel, walking class; `<10, 2, 5>` is the same code over 20 virtual stages)\n")
    isf = [1 if "DFMA" in t else 0 for _, t in ins]
    best, bi, s2 = -1, 0, sum(isf[:W])
    for i in range(len(ins) - W):
        if s2 > best:
            best, bi = s2, i
        s2 += isf[i + W] - isf[i]
    print(f"Densest FP64 window ({best} DFMA in {W} instructions): register-resident operands, one thread = one robot (no shuffles, no barriers, no shared memory):\n\n```")
    for a, t in ins[bi: bi + W]:
        print(f"/*{a:05x}*/ {t}")
    print("```\n")
    cs = [i for i, (_, t) in enumerate(ins) if ".EF" in t or "EVICT_FIRST" in t.upper() or ".CS" in t.upper()]
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for _, t in ins)
    print("Global accesses go to the lane-interleaved workspace (element i of lane l at `ws[i*32 + l]`: one 256-byte line pair per warp access); "
          "the stored factor and the row arrays use the streaming (evict-first) forms. Global/local memory opcodes of the kernel:\n\n```")
    for op, n in sorted(c.items(), key=lambda kv: -kv[1]):
        if op.split(".")[0] in ("LDG", "STG", "LD", "ST", "LDL", "STL", "LDC", "LDCU"):
            print(f"{n:6d}  {op}")
    print("```")
