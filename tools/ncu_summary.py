"""Summarise `ncu --set full` reports: one block of key metrics per captured kernel.
usage: python tools/ncu_summary.py report.ncu-rep [...]"""
import csv, subprocess, sys, io

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor_inst_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct2"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]


def main():
    for rep in sys.argv[1:]:
        if rep.endswith(".csv"):                 # a raw page exported on the GPU box (ncu -i rep --page raw --csv)
            out = open(rep).read()
        else:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(rep, "no data"); continue
        hdr, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(hdr)}
        print("== %s" % rep)
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            print("-- %s" % name[:110])
            parts = []
            for k, short in KEYS:
                if k in col and r[col[k]] not in ("", "n/a"):
                    parts.append("%s=%s%s" % (short, r[col[k]], (" " + units[col[k]]) if units[col[k]] not in ("", "%") else ""))
            print("   " + "; ".join(parts))
            stalls = []
            for n, i in col.items():
                if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") or \
                   (n.startswith("smsp__average_warp_latency_issue_stalled_") and n.endswith(".ratio")):
                    try:
                        stalls.append((float(r[i]), n.split("stalled_")[1].split("_per")[0].replace(".ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            if stalls:
                print("   top stalls: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:5]))


if __name__ == "__main__":
    main()
