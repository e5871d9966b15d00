"""Summarise an ncu raw CSV page: python ncu_summary.py raw.csv [metric-substring ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "sm__throughput.avg.pct",
                        "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate",
                        "smsp__issue_active.avg.pct", "issue_stalled", "sm__pipe_tensor", "launch__waves", "lts__t_bytes.sum ", "smsp__inst_executed.sum "]
for r in rows[2:]:
    print("=== kernel", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(k.strip() in h for k in keys):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if "issue_stalled" in h and v < 0.3:
                continue
            print("  %-95s %14.3f %s" % (h, v, units[i]))
