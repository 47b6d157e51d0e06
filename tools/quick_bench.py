"""Kernel-only timing of the fused forward on a few shapes (tuning aid; bench.py is the contract)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run(name, pairs, mode="ff", kind="smooth"):
    cfg = tcl.synth.CONFIGS[name]
    dt = torch.bfloat16 if cfg["dtype"] == "bf16" else torch.float32
    H, W = cfg["H"], cfg["W"]
    chunks = []
    for s in range(0, pairs, 32):
        n = min(32, pairs - s)
        ff, bf = tcl.synth.make_flows(n, H, W, seed=77 + s, max_shift=cfg["max_shift"], max_rot_deg=cfg["max_rot_deg"], device=dev)
        prev, cur = tcl.synth.make_frames(n, 3, H, W, seed=77 + s, kind=kind, device=dev, dtype=dt)
        chunks.append((ff, bf, prev, cur))
    ff, bf, prev, cur = (torch.cat([c[i] for c in chunks]) for i in range(4))
    del chunks
    if mode == "mask":
        m = tcl.fbcCheckTorch(ff, bf)
        fn = lambda: tcl.fused_forward(bf, prev, cur, mask=m, finalize=tcl.ops.FIN_MEAN)
        bpp = 36
    else:
        fn = lambda: tcl.fused_forward(bf, prev, cur, ff=ff)
        bpp = 40 if dt == torch.float32 else 28
    import ctypes
    lib = tcl._cabi.lib()
    lib.tclb200_debug_tile_stats(None, 1)
    fn()
    st = (ctypes.c_ulonglong * 2)()
    lib.tclb200_debug_tile_stats(st, 1)
    ms = timeit(fn)
    px = pairs * H * W
    print(f"{name:14s} pairs={pairs:4d} mode={mode:4s} {ms*1e3:9.1f} us  {px/ms/1e6:7.1f} Gpix/s  {px*bpp/ms/1e6:7.0f} GB/s  "
          f"{px*bpp/ms/1e6/6548.2*100:5.1f}% of measured peak  global/mixed tiles {st[0]}/{st[1]}", flush=True)


if __name__ == "__main__":
    run("sintel_full", 256)
    run("sintel_full", 256, kind="white")
    run("train_b16_256", 512, mode="mask")
    run("hd1080_window", 48)
    run("uhd4k_stress", 12)
