"""Small-launch latency: TMA kernel vs generic kernel at training-batch sizes (CUDA-graph replay, L2-rotating buffers)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
dev = torch.device("cuda:0")
lib = tcl._cabi.lib()

def bench(B, H, W, generic, mode, nbuf=12):
    bufs = []
    for i in range(nbuf):
        ff, bf = tcl.synth.make_flows(B, H, W, seed=100 + i, max_shift=24.0, max_rot_deg=6.0, device=dev)
        prev, cur = tcl.synth.make_frames(B, 3, H, W, seed=100 + i, device=dev)
        m = tcl.fbcCheckTorch(ff, bf)
        bufs.append((ff, bf, prev, cur, m))
    lib.tclb200_debug_force_generic(int(generic))
    def launch(i):
        ff, bf, prev, cur, m = bufs[i % nbuf]
        if mode == "mask":
            tcl.fused_forward(bf, prev, cur, mask=m, finalize=tcl.ops.FIN_MEAN)
        else:
            tcl.fused_forward(bf, prev, cur, ff=ff)
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for i in range(nbuf):
            launch(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(nbuf):
            launch(i)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    lib.tclb200_debug_force_generic(0)
    us = a.elapsed_time(b) / 20 / nbuf * 1e3
    print(f"B={B:3d} {H}x{W} mode={mode:4s} {'generic' if generic else 'tma    '}  {us:7.1f} us/launch  {B*H*W/us/1e3:6.1f} Gpix/s", flush=True)

for B in (4, 16, 64):
    for mode in ("mask", "ff"):
        for generic in (False, True):
            bench(B, 256, 256, generic, mode)
bench(1, 436, 1024, False, "ff"); bench(1, 436, 1024, True, "ff")
