"""Attribute executed warp-instructions to CUDA source lines.  usage: ncu_lines.py rep.ncu-rep units_per_kernel [top]"""
import collections
import csv
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
agg, samp, src = collections.Counter(), collections.Counter(), {}
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No") or not r[0].isdigit():
        continue
    try:
        ex = int(r[7])
    except ValueError:
        continue
    key = (cur, int(r[0]))
    agg[key] += ex
    samp[key] += int(r[4]) if r[4].isdigit() else 0
    src[key] = r[1].strip()
tot = sum(agg.values())
print(f"total executed warp-instructions {tot}  = {tot / units:.1f} per unit")
for key, v in agg.most_common(top):
    print(f"{v / units:7.1f}/unit {v / tot * 100:5.1f}%  stalls={samp[key]:5d}  {key[0]}:{key[1]}  {src[key][:105]}")
