#!/usr/bin/env python
"""Key metrics of one kernel from an ncu report:  python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__inst_executed.sum"]


def summary(path, table=False):
    """Full key-metric listing of every launch in the report, or (table=True) one compact line per launch."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = []
    if table:
        lines.append(f"{'kernel':44s} {'grid':>6s} {'time us':>9s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'dram %pk':>8s} {'L2 hit %':>8s} "
                     f"{'tensor %':>8s} {'regs':>5s} {'warps %':>7s}")

    def num(vals, k, scale=1.0):
        try:
            return float(vals[col[k]].replace(",", "")) * scale
        except (KeyError, ValueError):
            return float("nan")

    def to_base(vals, k):   # value in base units (bytes / seconds) from ncu's auto-scaled unit
        u = units[col[k]]
        f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
             "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
        return num(vals, k) * f

    for vals in rows[2:]:
        if len(vals) < len(hdr):
            continue
        if table:
            import re
            name = re.sub(r"\(.*", "", vals[col["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "").replace("dv3::<", "")
            grid = vals[col["Grid Size"]].strip("()").split(",")[0]
            lines.append(f"{name[:44]:44s} {grid:>6s} {to_base(vals, 'gpu__time_duration.sum') * 1e6:9.2f} "
                         f"{to_base(vals, 'dram__bytes_read.sum') / 1e6:10.2f} {to_base(vals, 'dram__bytes_write.sum') / 1e6:10.2f} "
                         f"{num(vals, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):8.1f} {num(vals, 'lts__t_sector_hit_rate.pct'):8.1f} "
                         f"{num(vals, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):8.2f} "
                         f"{num(vals, 'launch__registers_per_thread'):5.0f} {num(vals, 'sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f}")
        else:
            for k in KEYS:
                if k in col:
                    lines.append(f"{k:85s} {units[col[k]]:16s} {vals[col[k]]}")
            lines.append("")
    return lines


if __name__ == "__main__":
    print("\n".join(summary(sys.argv[1], table="--table" in sys.argv)))
