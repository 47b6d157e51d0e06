"""Sustained (power-capped) throughput of the fused computeTCL kernel: the number the bench reports.  usage: sustained.py [secs]"""
import os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
dev = torch.device("cuda:0")
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
cfg = tcl.synth.CONFIGS["sintel_full"]
pairs = 256
chunks = []
for s in range(0, pairs, 32):
    ff, bf = tcl.synth.make_flows(32, cfg["H"], cfg["W"], seed=77 + s, max_shift=32.0, max_rot_deg=3.0, device=dev)
    prev, cur = tcl.synth.make_frames(32, 3, cfg["H"], cfg["W"], seed=77 + s, device=dev)
    chunks.append((ff, bf, prev, cur))
ff, bf, prev, cur = (torch.cat([c[i] for c in chunks]) for i in range(4))
del chunks
px = pairs * cfg["H"] * cfg["W"]
fn = lambda: tcl.fused_forward(bf, prev, cur, ff=ff)
for _ in range(3):
    r = fn()
torch.cuda.synchronize()
check = float(r.total_val)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    fn()
b.record(); torch.cuda.synchronize()
burst = a.elapsed_time(b) / 20
samples = []
stop = threading.Event()
def sampler():
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        l = proc.stdout.readline()
        if l:
            samples.append((time.time(), l.strip()))
    proc.terminate()
threading.Thread(target=sampler, daemon=True).start()
t0 = time.time(); n = 0
a.record()
while time.time() - t0 < secs:
    for _ in range(50):
        fn()
    n += 50
    torch.cuda.synchronize()
b.record(); torch.cuda.synchronize()
t1 = time.time(); stop.set()
ms = a.elapsed_time(b) / n
sel = [s for (t, s) in samples if t0 + 1.0 < t < t1]
clk = [float(s.split(",")[0]) for s in sel]; pw = [float(s.split(",")[1]) for s in sel]
print(f"burst {px/burst/1e6:6.1f} Gpix/s | sustained {px/ms/1e6:6.1f} Gpix/s  sm {sum(clk)/max(len(clk),1):5.0f} MHz {sum(pw)/max(len(pw),1):5.0f} W  rmse {check:.9f}  [{os.path.basename(os.environ.get('TCL_B200_LIB','default'))}]", flush=True)
