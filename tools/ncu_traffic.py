"""Turn the ncu --set full capture of `python bench.py --no-extras ...` into profiles/r02_bench_traffic.json
(the `roofline.traffic` figure bench.py reports).  usage: ncu_traffic.py rep.ncu-rep workload pairs_per_launch"""
import csv
import json
import os
import subprocess
import sys

rep, workload, pairs = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
best = None
for vals in rows[2:]:
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    if "fused_forward_ws_kernel" not in m["Kernel Name"][0]:
        continue
    def val(k):
        v, u = m[k]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rec = dict(workload=workload, pairs_per_launch=pairs, kernel=m["Kernel Name"][0],
               dram_bytes_read=val("dram__bytes_read.sum"), dram_bytes_write=val("dram__bytes_write.sum"),
               gpu_time_us_under_ncu=float(m["gpu__time_duration.sum"][0].replace(",", "")),
               source=os.path.basename(rep))
    rec["dram_bytes_per_launch"] = rec["dram_bytes_read"] + rec["dram_bytes_write"]
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    rec["kernel_source_hash"] = bench.kernel_source_hash()   # bench.py reports the figure only while the kernel sources are these
    best = rec
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_bench_traffic.json")
json.dump(best, open(out, "w"), indent=1)
print(best)
