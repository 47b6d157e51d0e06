"""Clip mode (frames stored once) vs pairwise tensors: sustained throughput, clocks and power (3 s loops each)."""
import os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 257
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
cfg = tcl.synth.CONFIGS["sintel_full"]
H, W = cfg["H"], cfg["W"]
fl, fr = [], []
for s0 in range(0, T - 1, 32):
    n = min(32, T - 1 - s0)
    fl.append(tcl.synth.make_flows(n, H, W, seed=21 + s0, max_shift=32.0, max_rot_deg=3.0, device=dev))
for s0 in range(0, T, 32):
    fr.append(tcl.synth.make_frames(min(32, T - s0), 3, H, W, seed=21 + s0, device=dev)[0])
ff = torch.cat([c[0] for c in fl]); bf = torch.cat([c[1] for c in fl]); frames = torch.cat(fr)
prev, cur = frames[:-1].contiguous(), frames[1:].contiguous()
px = (T - 1) * H * W
samples = []
def sampler(stop):
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        l = proc.stdout.readline()
        if l:
            samples.append((time.time(), l.strip()))
    proc.terminate()
ops = {"pairwise tensors (40 B/px)": (lambda: tcl.temporal_error_per_pair(ff, bf, prev, cur), 40),
       "frames stored once (28 B/px)": (lambda: tcl.temporal_error_clip(frames, ff, bf), 28)}
only = os.environ.get("CLIP_MODE")
if only:
    ops = {k: v for k, v in ops.items() if k.startswith(only)}
stop = threading.Event()
threading.Thread(target=sampler, args=(stop,), daemon=True).start()
for name, (fn, bpp) in ops.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if secs <= 0:
        continue
    t0 = time.time(); n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.time() - t0 < secs:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    t1 = time.time()
    ms = a.elapsed_time(b) / n
    sel = [s for (t, s) in samples if t0 + 1.0 < t < t1]
    clk = [float(s.split(",")[0]) for s in sel]; pw = [float(s.split(",")[1]) for s in sel]
    print(f"{name:30s} {px/ms/1e6:7.1f} Gpix/s {px*bpp/ms/1e6:6.0f} GB/s algorithmic  sm {sum(clk)/max(len(clk),1):6.0f} MHz  {sum(pw)/max(len(pw),1):6.0f} W", flush=True)
    time.sleep(1.0)
stop.set()
