"""Sustained runs of several ops with nvidia-smi sampling: clocks / power under load (what burns the 1000 W cap?)."""
import os, subprocess, sys, threading, time
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
del chunks
m = tcl.fbcCheckTorch(ff, bf)
dst = torch.empty_like(prev)
px = pairs * cfg["H"] * cfg["W"]
samples = []
def sampler(stop):
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        l = proc.stdout.readline()
        if l:
            samples.append((time.time(), l.strip()))
    proc.terminate()
ops = {
    "fused computeTCL (ff)": (lambda: tcl.fused_forward(bf, prev, cur, ff=ff), 40),
    "fused loss (mask given)": (lambda: tcl.fused_forward(bf, prev, cur, mask=m, finalize=tcl.ops.FIN_MEAN), 36),
    "fbcCheckTorch only": (lambda: tcl.fbcCheckTorch(ff, bf), 20),
    "warp only": (lambda: tcl.warp(prev, bf), 32),
    "torch copy prev->dst": (lambda: dst.copy_(prev), 24),
}
stop = threading.Event()
th = threading.Thread(target=sampler, args=(stop,), daemon=True)
th.start()
for name, (fn, bpp) in ops.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.time()
    n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.time() - t0 < 3.0:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = a.elapsed_time(b) / n
    sel = [s for (t, s) in samples if t0 + 1.0 < t < t1]
    clk = [float(s.split(",")[0]) for s in sel]
    pw = [float(s.split(",")[1]) for s in sel]
    print(f"{name:28s} {px/ms/1e6:7.1f} Gpix/s {px*bpp/ms/1e6:6.0f} GB/s  sm {sum(clk)/max(len(clk),1):6.0f} MHz  {sum(pw)/max(len(pw),1):6.0f} W  ({len(sel)} samples)", flush=True)
    time.sleep(1.0)
stop.set()
