"""Aggregate warp-stall samples and executed instructions per CUDA source line of an ncu report
(captured with --import-source on): python bench/ncu_lines.py file.ncu-rep [top]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, agg, tot_s, tot_i = None, [], 0, 0
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        print("==", r[1][:100])
    if len(r) > 8 and r[0].isdigit():
        try:
            s, ins = int(r[6]), int(r[7])
        except ValueError:
            continue
        agg.append((s, ins, cur, r[0], r[1][:100]))
        tot_s += s
        tot_i += ins
agg.sort(reverse=True)
print("total samples", tot_s, "warp instructions", tot_i)
for s, ins, f, ln, src in agg[:top]:
    print(f"{100 * s / tot_s:5.1f}% smp {100 * ins / tot_i:5.1f}% ins  {f}:{ln}  {src}")
