"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line.
usage: python tools/ncu_lines.py export.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or not r:
        continue
    if r[0] != "":
        try:
            samples = int(r[4]); inst = int(r[7])
        except Exception:
            continue
        agg.append((samples, inst, cur_file, r[0], r[1].strip()[:110]))
tot = sum(a[0] for a in agg)
print("total samples", tot)
for s, i, f, ln, src in sorted(agg, reverse=True)[:top]:
    print("%6d %5.1f%% inst=%9d %s:%s  %s" % (s, 100.0 * s / max(tot, 1), i, f, ln, src))
