import os, sys, torch
sys.path.insert(0, "/root/repo")
import tcl_b200 as tcl
dev = torch.device("cuda:0")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (B, C, H, W) in [(16, 64, 64, 64), (16, 128, 64, 64), (16, 48, 128, 128), (4, 2, 436, 1024)]:
    ff, bf = tcl.synth.make_flows(B, H, W, seed=1, max_shift=6.0, device=dev)
    x = torch.randn(B, C, H, W, device=dev)
    ms = timeit(lambda: tcl.fs_warp(x, bf))
    by = B * H * W * (8 + 8 * C)
    print(f"fs_warp B={B} C={C} {H}x{W}: {ms*1e3:7.1f} us  {by/ms/1e6:6.0f} GB/s")
    if C % 3 == 0:
        xg = x.view(B * C // 3, 3, H, W)
        fg = bf.repeat_interleave(C // 3, dim=0)
        ms2 = timeit(lambda: tcl.fs_warp(x.view(B * C // 3, 3, H, W), bf.repeat_interleave(C // 3, dim=0)))
        print(f"   as {B*C//3} groups of 3 channels (flow repeated): {ms2*1e3:7.1f} us")
