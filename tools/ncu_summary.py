"""Dev tool: key metrics of every kernel in an ncu report -> CSV (profiles/*.csv).

usage: python tools/ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        w.writerow(["Kernel Name", "", name])
        for i, k in enumerate(hdr):
            if k in KEYS or (k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and "not_issued" not in k):
                w.writerow([k, units[i], r[i]])
print("wrote", out)
