"""Aggregate an `ncu --page source --print-source sass --csv` export by SASS opcode: stall samples, executed warp instructions and
shared-memory wavefronts per unit of work.  usage: python tools/ncu_opcodes.py export.csv <units (e.g. matrices)> [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    op = re.sub(r'^@!?U?P\d+\s+', '', r[col['Source']].strip()).split()[0]
    a = agg[op]
    a[0] += int(r[col['# Samples']])
    a[1] += int(r[col['Instructions Executed']])
    a[2] += int(r[col['L1 Wavefronts Shared']] or 0)
    a[3] += 1
tot = sum(a[0] for a in agg.values())
print("%s: warp instructions per unit %.1f, shared-memory wavefronts per unit %.1f" %
      (rows[0][1][:60], sum(a[1] for a in agg.values()) / N, sum(a[2] for a in agg.values()) / N))
for op, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%-18s stall samples %5.1f%%  per unit %7.1f  smem wavefronts %7.1f  static %d" % (op, 100.0 * a[0] / tot, a[1] / N, a[2] / N, a[3]))
