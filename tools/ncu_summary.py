#!/usr/bin/env python
"""Summarise `ncu --set full` reports (one line block per kernel launch): duration, DRAM bytes, achieved DRAM rate,
occupancy, registers, the stall reasons that matter for streaming kernels.  Usage: ncu_summary.py a.ncu-rep [b.ncu-rep]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__bytes.sum.per_second", "dram rate"), ("launch__registers_per_thread", "registers"),
        ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("launch__grid_size", "grid")]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        H, U = rows[0], rows[1]
        print("# %s" % path)
        for r in rows[2:]:
            name = r[H.index("Kernel Name")].replace("mgb::", "").split("(")[0]
            print(name)
            for k, label in KEYS:
                if k in H:
                    i = H.index(k)
                    print("    %-34s %s %s" % (label, r[i], U[i]))


if __name__ == "__main__":
    main()
