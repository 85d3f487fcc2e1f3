"""Dev tool: per-kernel summaries of an ncu report for profiles/.

    ncu -i report.ncu-rep --page raw --csv > raw.csv
    python tests/dev/ncu_summary.py raw.csv profiles/PREFIX      # writes PREFIX_<n>_<kernel>_summary.csv

Keeps the metrics DESIGN.md quotes (time, DRAM bytes, launch shape, occupancy limits, pipe utilisation, issue-stall
breakdown).  bench.py's roofline.traffic comes from profiles/traffic.json, which is filled from these files by hand
together with the observation count of the captured launch."""
import csv, re, sys

KEEP = re.compile(r"^(Kernel Name|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"gpu__time_duration\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|launch__(block_size|grid_size|"
                  r"occupancy_limit_registers|occupancy_limit_shared_mem|registers_per_thread|shared_mem_per_block_dynamic|cluster_x)|"
                  r"sm__inst_executed_pipe_(alu|fp64|lsu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active)$")

def main(raw, prefix):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        short = re.sub(r"[^A-Za-z0-9]+", "_", name.split("(")[0].replace("void ", "")).strip("_")
        path = f"{prefix}_{n}_{short}_summary.csv"
        with open(path, "w", newline="") as f:
            w = csv.writer(f)
            for h, u, v in sorted(zip(hdr, units, r)):
                if KEEP.match(h):
                    w.writerow([h, u, v])
        print(path)

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
