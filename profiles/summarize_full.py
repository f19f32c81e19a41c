#!/usr/bin/env python
"""Reduce `ncu -i X.ncu-rep --page raw --csv` (all metrics of an `ncu --set full` capture) to the columns the roofline
discussion uses. usage: python profiles/summarize_full.py gpurun_out/prof_a_raw.csv [more.csv ...] > profiles/rNN_full.txt"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("launch__grid_size", "grid"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_lsb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_ssb"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar")]


def main(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        cols = [(c, n) for c, n in COLS if c in idx]
        print("# %s: %d launches" % (path, len(rows) - 2))
        print("%-34s " % "kernel" + " ".join("%12s" % n for _, n in cols))
        for r in rows[2:]:
            cells = []
            for c, _ in cols:
                v, u = r[idx[c]], units[idx[c]]
                try:
                    cells.append("%12s" % ("%.4g%s" % (float(v.replace(",", "")), {"Gbyte": "GB", "Mbyte": "MB", "Kbyte": "KB", "byte": "B", "us": "us", "ms": "ms", "%": ""}.get(u, ""))))
                except ValueError:
                    cells.append("%12s" % v[:12])
            print("%-34s " % r[idx["Kernel Name"]].split("(")[0].replace("pr::", "")[:34] + " ".join(cells))


if __name__ == "__main__":
    main(sys.argv[1:])
