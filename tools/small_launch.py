"""Small-launch latency: direct kernel vs TMA pipeline vs generic kernel at training-batch sizes (CUDA-graph replay, L2-rotating buffers)."""
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
    name = {0: "auto   ", 1: "generic", 2: "tma    ", 3: "direct "}[int(generic)]
    print(f"B={B:3d} {H}x{W} mode={mode:4s} {name}  {us:7.1f} us/launch  {B*H*W/us/1e3:6.1f} Gpix/s", flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "one":     # for ncu: a few plain launches of the training shape on the direct kernel
    ff, bf = tcl.synth.make_flows(16, 256, 256, seed=100, max_shift=24.0, max_rot_deg=6.0, device=dev)
    prev, cur = tcl.synth.make_frames(16, 3, 256, 256, seed=100, device=dev)
    m = tcl.fbcCheckTorch(ff, bf)
    for _ in range(3):
        tcl.fused_forward(bf, prev, cur, mask=m, finalize=tcl.ops.FIN_MEAN)
    torch.cuda.synchronize()
    sys.exit(0)
for B in (4, 16, 64, 128):
    for mode in ("mask", "ff"):
        for force in ((3, 2, 1) if mode == "mask" else (2, 1)):
            if B == 128 and force == 1:
                continue
            bench(B, 256, 256, force, mode, nbuf=12 if B <= 16 else 4)
for B in (1, 8, 24):
    for force in (3, 2):
        bench(B, 436, 1024, force, "mask", nbuf=8 if B == 1 else 3)
bench(1, 436, 1024, 2, "ff"); bench(1, 436, 1024, 1, "ff")
