"""Executed-instruction and stall-sample histogram by SASS opcode from `ncu --page source --csv` output.
usage: ncu -i rep.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv; python tools/ncu_opcodes.py src.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
st = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
hdr = rows[st[0] + 1]
isrc, iex, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[st[0] + 2:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?', r[isrc])
    if not m:
        continue
    op = m.group(2)
    key = op + (m.group(3) or '') if op in ('F2F', 'I2F', 'F2I', 'MUFU', 'BAR') else op
    ops[key] += int(r[iex])
    samp[key] += int(r[isamp])
    tot += int(r[iex])
ts = sum(samp.values())
print('total executed', tot, 'samples', ts)
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print('%-18s %6.2f%% inst  %6.2f%% samples' % (k, 100 * v / tot, 100 * samp[k] / ts))
