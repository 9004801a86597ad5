"""Prints the handful of ncu raw-page metrics the kernel notes quote.  Usage: ncu_brief.py report.ncu-rep [kernel-regex]"""
import csv, io, subprocess, sys
args = ["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"]
if len(sys.argv) > 2:
    args += ["--kernel-name", "regex:" + sys.argv[2]]
out = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    print("#", r[h.index("Kernel Name")] if "Kernel Name" in h else "")
    for i, n in enumerate(h):
        if n in WANT:
            print(f"{n:100s} {units[i]:18s} {r[i]}")
        elif "issue_stalled" in n and n.endswith("per_issue_active.ratio"):
            try:
                if float(r[i]) >= 0.05:
                    print(f"{n:100s} {units[i]:18s} {r[i]}")
            except ValueError:
                pass
