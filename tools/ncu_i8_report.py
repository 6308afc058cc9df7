"""profiles/r01_ncu_i8.md + traffic.json entries from the two `tools/profile_i8.sh` captures (one sweep each)."""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                      os.path.join(ROOT, "gpurun_out", "prof_r01_i8_c2.ncu-rep"),
                      os.path.join(ROOT, "gpurun_out", "prof_r01_i8_c3.ncu-rep")], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", "r01_ncu_i8_raw_summary.txt"), "w").write(raw)
def tobytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
rows, out, table = 151552, {}, []
for rep, q in zip(raw.split("== ")[1:], (16, 32)):
    for blk in rep.split("-- ")[1:]:
        name = blk.split("\n")[0]
        short = re.sub(r"^(void )?(unnamed>::)?", "", name).split("(")[0].split("<")[0]
        rd, wr = re.search(r"dram_read=([0-9.]+) (\w+)", blk), re.search(r"dram_write=([0-9.]+) (\w+)", blk)
        dur = float(re.search(r"duration=([0-9.]+)", blk).group(1))
        b = tobytes(*rd.groups()) + tobytes(*wr.groups())
        key = {"zsolve_tpm_kernel": "zsolve", "zsolve_blocked_kernel": "zsolve", "zstep_i8_kernel": "zstep_i8",
               "zstep_dmma_kernel": "zstep_eta", "digitize_kernel": "digitize", "stats_i8_kernel": "stats_i8",
               "stats_dmma_kernel": "stats_x"}.get(short)
        g = lambda k: (re.search(k + r"=([0-9.]+)", blk) or [0, "0"])[1]
        st = re.search(r"top stalls: (.*)", blk)
        table.append((q, short, dur, b / 1e6, b / rows, g("tensor_pipe_pct"), g("fp64_pipe_pct"), g("dram_pct"), st.group(1) if st else ""))
        if key:
            out["%s%d_dram_bytes_per_row" % (key, q)] = b / rows
p = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(p)); t.update(out)
json.dump(t, open(p, "w"), indent=1)
tot = {q: sum(r[2] for r in table if r[0] == q) for q in (16, 32)}
md = ["# ncu evidence, INT8-path sweep (round 1, final kernels)", "",
      "`tools/profile_i8.sh`: `ncu --profile-from-start off --set full --clock-control none --import-source on` around ONE steady-state sweep",
      "(`tools/profile_sweep.py`, N = 151,552 rows per launch) after the same command had exited 0 without ncu; raw numbers in",
      "`r01_ncu_i8_raw_summary.txt`, launch list of the bench command in `r01_launches_i8.csv`.  Durations under ncu are cold-cache and",
      "serialised: use them for the SHARE of each kernel; the absolute times are in `r01_bench_c2_1gpu_i8.json` (CUDA events).", "",
      "| q (D) | kernel | duration us | share % | DRAM MB | DRAM B/row | tensor pipe % | FP64 pipe % | DRAM % of peak | top stalls |",
      "|---|---|---|---|---|---|---|---|---|---|"]
for q, short, dur, mb, bpr, tp, fp, dp, st in table:
    md.append("| %d (%d) | %s | %.1f | %.1f | %.1f | %.0f | %s | %s | %s | %s |" % (q, 256 if q == 16 else 1024, short, dur, 100 * dur / tot[q], mb, bpr, tp[:5], fp[:5], dp[:5], st[:60]))
md += ["", "Reading: at the C2 shape (q = 16, D = 256) the sweep is 13 launches and the seven streaming kernels take ~95 % of it.  `zsolve_tpm` (K2) is",
       "the largest single kernel, latency-bound on its 5 warps per SM (the 32-matrix batch of a warp fills 40 KB of shared memory).  `zstep_i8` /",
       "`stats_i8` keep the INT8 tensor pipe 30-60 % busy: after the issue loops were fixed (elected lane, descriptors by adds) the limit is the",
       "number of digit-tile stages that fit next to the resident mask block against the L2 -> SM latency.  The two FP64 tensor kernels that remain",
       "(`zstep_dmma<ETA>`, `stats_dmma<XO>`) are balanced between the DMMA pipe and DRAM at q = 16 (4 flop per byte) and DMMA-bound (~75 %) at",
       "q = 32.  `digitize` streams at ~60-70 % of the DRAM peak.  DRAM bytes per row are at or below the algorithmic bytes of DESIGN.md section 5a",
       "for every kernel (no re-reads; part of each kernel's output is still in L2 when it ends)."]
open(os.path.join(ROOT, "profiles", "r01_ncu_i8.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[9:36]))
