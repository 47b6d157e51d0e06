"""Pipeline timing of the warp-specialised kernel (needs a -DTCL_TRACE build: tools/sweep_build.py trace).

Prints, per local tile of a few CTAs: when the source boxes were requested, how long the request took as seen by
consumer warp 0 (request -> ready), how long that warp waited for it, and the tile's compute time.
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl  # noqa: E402

dev = torch.device("cuda:0")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = tcl.synth.CONFIGS["sintel_full"]
ff, bf = tcl.synth.make_flows(pairs, cfg["H"], cfg["W"], seed=3234, max_shift=cfg["max_shift"], max_rot_deg=cfg["max_rot_deg"], device=dev)
prev, cur = tcl.synth.make_frames(pairs, 3, cfg["H"], cfg["W"], seed=3234, kind="smooth", device=dev)
for _ in range(3):
    tcl.fused_forward(bf, prev, cur, ff=ff)
torch.cuda.synchronize()
lib = ctypes.CDLL(tcl._cabi.LIB_PATH)
buf = np.zeros((160, 64, 4), dtype=np.uint64)
lib.tclb200_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
assert lib.tclb200_debug_trace(buf.ctypes.data, buf.nbytes) == 0
t = buf.astype(np.int64)
n = min(48, (pairs * 16 * 14) // 148)
for cta in (0, 37, 147):
    r = t[cta, :n]
    base = r[0, 0]
    print(f"CTA {cta}: tile  req@us  req->ready(us)  waited(us)  compute(us)  req-lead-before-need(us)")
    for k in range(2, n, 3):
        print(f"   {k:3d}  {(r[k,0]-base)/1e3:8.2f}  {(r[k,2]-r[k,0])/1e3:8.2f}  {(r[k,2]-r[k,1])/1e3:8.2f}  {(r[k,3]-r[k,2])/1e3:8.2f}  {(r[k,1]-r[k,0])/1e3:8.2f}")
a = t[:148, 4:n]
print("mean over CTAs/tiles: req->ready %.2f us, waited %.2f us, compute %.2f us, tile period %.2f us" % (
    ((a[:, :, 2] - a[:, :, 0]).mean() / 1e3), ((a[:, :, 2] - a[:, :, 1]).mean() / 1e3), ((a[:, :, 3] - a[:, :, 2]).mean() / 1e3),
    ((a[:, -1, 3] - a[:, 0, 3]).mean() / 1e3 / (a.shape[1] - 1))))
