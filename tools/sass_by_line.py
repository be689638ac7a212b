"""Attribute SASS instructions of one kernel to source lines (nvdisasm -g output with '//## File "..", line N' markers).

usage: nvdisasm -g -c <cubin> | python tools/sass_by_line.py <kernel-name-substring> [opcode-regex]
Prints, per source line, the count of instructions whose opcode matches the regex (default: all)."""
import collections
import re
import sys

want = sys.argv[1]
opre = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
cur_fn, cur_line, in_fn = None, None, False
counts = collections.Counter()
total = 0
for ln in sys.stdin:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        in_fn = want in m.group(1)
        continue
    if ln.startswith("\t.section") or ln.startswith(".section"):
        in_fn = False if ".text." not in ln else in_fn
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        if opre is None or opre.search(op):
            counts[cur_line] += 1
            total += 1
print("total", total)
for (f, l), c in sorted(counts.items(), key=lambda kv: (kv[0] or ("", 0))):
    print(f"{f}:{l}\t{c}")
