"""Sustained run of the fused kernel with nvidia-smi sampling: clocks / power under load."""
import os, subprocess, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
dev = torch.device("cuda:0")
cfg = tcl.synth.CONFIGS["sintel_full"]
pairs = 256
chunks = []
for s in range(0, pairs, 32):
    ff, bf = tcl.synth.make_flows(32, cfg["H"], cfg["W"], seed=77 + s, max_shift=32.0, max_rot_deg=3.0, device=dev)
    prev, cur = tcl.synth.make_frames(32, 3, cfg["H"], cfg["W"], seed=77 + s, device=dev)
    chunks.append((ff, bf, prev, cur))
ff, bf, prev, cur = (torch.cat([c[i] for c in chunks]) for i in range(4))
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
px = pairs * cfg["H"] * cfg["W"]
for rep in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(500):
        tcl.fused_forward(bf, prev, cur, ff=ff)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 500
    print(f"block {rep}: {ms*1e3:.1f} us/launch  {px/ms/1e6:.1f} Gpix/s", flush=True)
proc.terminate()
out = proc.stdout.read().splitlines()
print("nvidia-smi samples (sm MHz, W, C, power_cap, reasons):")
for l in out[::4]:
    print("  ", l)
