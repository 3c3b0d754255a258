"""tools/ncu_summary.py REPORT.ncu-rep [regex] -- print the metrics that matter for an
HBM-bound kernel from an ncu report (one column per captured launch)."""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    r"gpu__time_duration\.sum", r"dram__bytes_read\.sum$", r"dram__bytes_write\.sum$",
    r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed", r"lts__t_sector_hit_rate\.pct",
    r"lts__t_bytes\.sum$", r"l1tex__t_bytes\.sum$",
    r"launch__registers_per_thread", r"launch__occupancy_limit", r"launch__waves_per_multiprocessor",
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active", r"sm__throughput\.avg\.pct",
    r"smsp__issue_active\.avg\.pct", r"smsp__inst_executed\.sum$", r"smsp__inst_executed\.avg\.per_cycle_active",
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$",
    r"smsp__average_warps_issue_stalled_.*_per_issue_active", r"smsp__warps_eligible\.avg\.per_cycle_active",
    r"launch__shared_mem_per_block", r"sm__maximum_warps_per_active_cycle_pct", r"launch__grid_size", r"launch__block_size",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2] if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("kernels:", [r[hdr.index("Kernel Name")][:70] for r in data])
    pats = [re.compile(k) for k in KEYS] + ([re.compile(extra)] if extra else [])
    for i, h in enumerate(hdr):
        if any(p.search(h) for p in pats):
            vals = [r[i] for r in data]
            if all(v in ("0", "0.00", "") for v in vals) and "stalled" in h:
                continue
            print("%-95s %-12s %s" % (h[-95:], units[i], "  ".join(vals)))


if __name__ == "__main__":
    main()
