"""Stall samples and executed warp instructions of the LDS smoother kernel per PHASE (source-line ranges of kernels_lds.cu), from an
`ncu --page source --print-source cuda,sass --csv` export.  usage: python tools/ncu_lds_phases.py export.csv <sequences>"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
B = float(sys.argv[2])
cur, hdr, agg = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or r[0] == "":
        continue
    try:
        agg.append((cur, int(r[0]), int(r[4]), int(r[7])))
    except ValueError:
        pass
tot = sum(a[2] for a in agg)
print("stall samples %d, warp instructions per sequence %.0f" % (tot, sum(a[3] for a in agg) / B))
src = open(os.path.join(ROOT, "pyvb_b200", "csrc", "kernels_lds.cu")).read().split("\n")


def find(s):
    return next(i + 1 for i, l in enumerate(src) if s in l)


marks = sorted([("load", find("---- load the sequence")), ("precisions", find("---- expected precisions")),
                ("gains", find("---- smoother gains")), ("boundary step (lambda)", find("auto step = ")),
                ("serial sweeps", find("if (!SCAN) {")), ("A: u_t in parallel", find("A. u_t =")),
                ("B: K^L (binary powering)", find("B. the recurrence matrix")), ("chunk body (passes 1 + 2)", find("auto chunk = ")),
                ("pass 1 store + carries", find("chunk(x, false)") - 3), ("pass 2 start", find("pass 2 from the true chunk")),
                ("sweep calls", find("sweep(true);") - 1), ("statistics", find("---- sufficient statistics")),
                ("parameters", find("---- parameters: lane")), ("store", find("---- write back")), ("end", len(src))], key=lambda x: x[1])
for (n, a), (_, b) in zip(marks, marks[1:]):
    s = sum(x[2] for x in agg if x[0] == "kernels_lds.cu" and a <= x[1] < b)
    i = sum(x[3] for x in agg if x[0] == "kernels_lds.cu" and a <= x[1] < b)
    print("%-28s lines %3d-%3d  stall samples %5.1f%%  instructions per sequence %7.0f" % (n, a, b, 100.0 * s / tot, i / B))
oth = {f: round(100.0 * sum(x[2] for x in agg if x[0] == f) / tot, 1) for f in set(x[0] for x in agg) if f != "kernels_lds.cu"}
print("inlined from other files (stall samples %):", oth)
