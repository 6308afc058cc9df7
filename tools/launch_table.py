"""Print kernel name / duration of the last `n` launches of an ncu --metrics gpu__time_duration.sum CSV log."""
import csv, sys
f, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows[1:][-n:]:
    print("%-70s %12s %s" % (r[ki][:70], r[vi], r[ui]))
