"""profiles/r02_ncu.md + traffic.json entries from the raw ncu pages of `tools/profile_r02.sh`
(one steady-state sweep at config 2, N = 1M, and at the config-3 shard, N = 1.25M: the BENCH sizes)."""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
CFG = [("c2", 1000000, 256, 16), ("c3", 1250000, 1024, 32)]
raw = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py")] +
                     [os.path.join(G, "r02_prof_%s_raw.csv" % c[0]) for c in CFG], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", "r02_ncu_raw_summary.txt"), "w").write(raw)


def tobytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


KEY = {"zsolve_gj_kernel": "zsolve", "zsolve_tpm_kernel": "zsolve", "zsolve_blocked_kernel": "zsolve", "zstep_i8_kernel": "zstep_i8", "zstep_dmma_kernel": "zstep_eta",
       "digitize_kernel": "digitize", "stats_i8_kernel": "stats_i8", "stats_dmma_kernel": "stats_x"}
out, table = {}, []
for rep, (name, rows, D, q) in zip(raw.split("== ")[1:], CFG):
    for blk in rep.split("-- ")[1:]:
        kname = blk.split("\n")[0]
        short = re.sub(r"^(void )?(unnamed>::)?", "", kname).split("(")[0].split("<")[0]
        rd, wr = re.search(r"dram_read=([0-9.]+) (\w+)", blk), re.search(r"dram_write=([0-9.]+) (\w+)", blk)
        dm = re.search(r"duration=([0-9.]+) (\w+)", blk)
        dur = float(dm.group(1)) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(dm.group(2), 1.0)
        b = tobytes(*rd.groups()) + tobytes(*wr.groups())
        g = lambda k: (re.search(k + r"=([0-9.]+)", blk) or [0, "0"])[1]
        st = re.search(r"top stalls: (.*)", blk)
        table.append((name, q, D, short, dur, b / 1e6, b / rows, g("tensor_pipe_pct"), g("fp64_pipe_pct"), g("dram_pct"), g("issue_active_pct"),
                      st.group(1) if st else ""))
        if short in KEY:      # (the conditional fall-back launches of the same kernels exit at once: keep the launch that did the work)
            k = "%s%d_dram_bytes_per_row" % (KEY[short], q)
            out[k] = max(out.get(k, 0.0), b / rows)
p = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(p))
t.update(out)
t["note_r02"] = ("round 2: INT8-path kernels re-measured at the BENCH sizes (ncu --set full, one steady-state sweep at N = 1,000,000 "
                 "(q = 16) / N = 1,250,000 (q = 32), profiles/r02_ncu_raw_summary.txt); dram read+write bytes per row")
json.dump(t, open(p, "w"), indent=1)
md = ["# ncu evidence, round 2: one sweep of the default (INT8) path at the bench sizes", "",
      "`tools/profile_r02.sh`: `ncu --profile-from-start off --set full --clock-control none` around ONE steady-state sweep",
      "(`tools/profile_sweep.py`; config 2: N = 1,000,000, D = 256, q = 16, 20 % missing; config-3 shard: N = 1,250,000, D = 1024, q = 32, 30 %)",
      "after the same command had exited 0 without ncu.  The reports (35 / 50 MB) stayed on the GPU box; their raw pages were exported to CSV",
      "there and are summarised in `r02_ncu_raw_summary.txt`.  Launch list of the bench command: `r02_launches_bench.csv`.  Durations under ncu are",
      "cold-cache and serialised: use them for the SHARE of each kernel; the event-timed numbers are in `r02_bench_1gpu.json`.", "",
      "| config | kernel | duration us | share % | DRAM MB | DRAM B/row | tensor pipe % | FP64 pipe % | DRAM % of peak | issue % | top stalls |",
      "|---|---|---|---|---|---|---|---|---|---|---|"]
tot = {c[0]: sum(r[4] for r in table if r[0] == c[0]) for c in CFG}
sums = {c[0]: sum(r[6] for r in table if r[0] == c[0]) for c in CFG}
for name, q, D, short, dur, mb, bpr, tp, fp, dp, ia, st in table:
    md.append("| %s (D %d, q %d) | %s | %.1f | %.1f | %.1f | %.0f | %s | %s | %s | %s | %s |" % (name, D, q, short, dur, 100 * dur / tot[name], mb, bpr,
              tp[:5], fp[:5], dp[:5], ia[:5], st[:70]))
md += ["", "Sweep totals (sum over the kernels of one sweep): " + "; ".join(
    "%s: %.0f us under ncu, %.0f DRAM bytes per row (algorithmic 8D + 8q + 8P = %d)" % (c[0], tot[c[0]], sums[c[0]], 8 * c[2] + 8 * c[3] + 4 * c[3] * (c[3] + 1))
    for c in CFG) + "."]
open(os.path.join(ROOT, "profiles", "r02_ncu.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[8:]))
