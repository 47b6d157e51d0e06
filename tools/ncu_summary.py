"""Summarise an .ncu-rep (raw + source pages) into a few lines.  usage: ncu_summary.py file.ncu-rep [out.txt]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "inst_executed",
        "smsp__sass_inst_executed_op_shared_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]
out = []
for k in keys:
    if k in m:
        out.append(f"{k:75s} {m[k][0]} {m[k][1]}")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = float(m[h][0].replace(",", ""))
        if v >= 0.1:
            out.append(f"{h:75s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))[2:]
by, samp, tot = collections.Counter(), collections.Counter(), 0
for r in srows:
    op = r[1].strip().split()
    if not op:
        continue
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    by[o] += int(r[5]); samp[o] += int(r[2]); tot += int(r[5])
out.append(f"warp-instructions executed: {tot}")
out.append("opcode mix (share of executed warp-instructions / stall samples): " +
           ", ".join(f"{o} {c / tot * 100:.1f}%/{samp[o]}" for o, c in by.most_common(16)))
out.append("top stall sites:")
for r in sorted(srows, key=lambda r: -int(r[2]))[:14]:
    out.append(f"  samples={r[2]:>6s} executed={r[5]:>9s}  {r[1].strip()[:80]}")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
