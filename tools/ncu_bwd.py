"""ncu target: fused training-loss backward on a Sintel-shape batch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
dev = torch.device("cuda:0")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = tcl.synth.CONFIGS["sintel_full"]
ff, bf = tcl.synth.make_flows(pairs, cfg["H"], cfg["W"], seed=5, max_shift=32.0, max_rot_deg=3.0, device=dev)
prev, cur = tcl.synth.make_frames(pairs, 3, cfg["H"], cfg["W"], seed=5, device=dev)
m = tcl.fbcCheckTorch(ff, bf)
p2 = prev.clone().requires_grad_(True)
c2 = cur.clone().requires_grad_(True)
for _ in range(3):
    p2.grad = None; c2.grad = None
    tcl.temporal_loss(m, c2, p2, bf).backward()
torch.cuda.synchronize()
print("ok", float(p2.grad.abs().sum()))
