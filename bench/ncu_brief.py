"""Print the handful of metrics we quote from an ncu report: python bench/ncu_brief.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__registers_per_thread", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for v in rows[2:]:
    print("==", v[h.index("Kernel Name")] if "Kernel Name" in h else "")
    for n, un, x in zip(h, u, v):
        if n in WANT or ("stalled" in n and "per_issue_active" in n and float(x or 0) > 0.3):
            print(f"  {n} = {x} {un}")
