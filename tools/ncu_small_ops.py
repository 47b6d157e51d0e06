"""ncu targets for the specialised single-purpose kernels: `python tools/ncu_small_ops.py gradient|mob|outputs` runs one call
(after a warm-up) of gradient(bf[:,0]) / fbcCheckTorch_mob / the fused error with warp+mask outputs on 128 Sintel-shape pairs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl

what = sys.argv[1]
d = torch.device("cuda:0")
cfg = tcl.synth.CONFIGS["sintel_full"]
parts = [tcl.synth.make_flows(32, cfg["H"], cfg["W"], seed=40 + i, max_shift=32.0, max_rot_deg=3.0, device=d) for i in range(4)]
ff, bf = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
fn = {"gradient": lambda: tcl.gradient(bf[:, 0]), "mob": lambda: tcl.fbcCheckTorch_mob(ff, bf)}.get(what)
if fn is None:
    fr = [tcl.synth.make_frames(32, 3, cfg["H"], cfg["W"], seed=40 + i, device=d) for i in range(4)]
    prev, cur = torch.cat([f[0] for f in fr]), torch.cat([f[1] for f in fr])
    fn = lambda: tcl.fused_forward(bf, prev, cur, ff=ff, want_warp=True, want_mask=True)
fn(); fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
