"""One process, a few launches of the fused computeTCL kernel on Sintel-shape pairs: the target of ncu captures.
usage: prof_hot.py [pairs] [launches]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl  # noqa: E402

dev = torch.device("cuda:0")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = tcl.synth.CONFIGS["sintel_full"]
chunks = []
for s in range(0, pairs, 32):
    n = min(32, pairs - s)
    ff, bf = tcl.synth.make_flows(n, cfg["H"], cfg["W"], seed=77 + s, max_shift=cfg["max_shift"], max_rot_deg=cfg["max_rot_deg"], device=dev)
    prev, cur = tcl.synth.make_frames(n, 3, cfg["H"], cfg["W"], seed=77 + s, device=dev)
    chunks.append((ff, bf, prev, cur))
ff, bf, prev, cur = (torch.cat([c[i] for c in chunks]) for i in range(4))
del chunks
for _ in range(launches):
    r = tcl.fused_forward(bf, prev, cur, ff=ff)
torch.cuda.synchronize()
print("rmse", float(r.total_val))
