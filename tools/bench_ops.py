"""Kernel-level timing of every public op on the BASELINE shapes (tuning aid; bench.py is the contract)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6548.2


def timeit(fn, n=20, warm=3):
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


def report(name, ms, px, bpp):
    print(f"{name:44s} {ms*1e3:9.1f} us  {px/ms/1e6:7.1f} Gpix/s  {px*bpp/ms/1e6:7.0f} GB/s  {px*bpp/ms/1e6/PEAK*100:5.1f}% of measured peak", flush=True)


def run(cfg_name, pairs):
    cfg = tcl.synth.CONFIGS[cfg_name]
    H, W = cfg["H"], cfg["W"]
    chunks = []
    for s in range(0, pairs, 32):
        n = min(32, pairs - s)
        ff, bf = tcl.synth.make_flows(n, H, W, seed=11 + s, max_shift=cfg["max_shift"], max_rot_deg=cfg["max_rot_deg"], device=dev)
        prev, cur = tcl.synth.make_frames(n, 3, H, W, seed=11 + s, device=dev)
        chunks.append((ff, bf, prev, cur))
    ff, bf, prev, cur = (torch.cat([c[i] for c in chunks]) for i in range(4))
    del chunks
    px = pairs * H * W
    tag = f"{cfg_name}[{pairs}] "
    report(tag + "warp fp32", timeit(lambda: tcl.warp(prev, bf)), px, 32)
    report(tag + "fs_warp fp32", timeit(lambda: tcl.fs_warp(prev, bf)), px, 32)
    report(tag + "fbcCheckTorch", timeit(lambda: tcl.fbcCheckTorch(ff, bf)), px, 20)
    report(tag + "fbcCheckTorch_mob", timeit(lambda: tcl.fbcCheckTorch_mob(None, bf)), px, 12)
    report(tag + "gradient(u)", timeit(lambda: tcl.gradient(bf[:, 0])), px, 12)
    m = tcl.fbcCheckTorch(ff, bf)
    report(tag + "temporal_error (fused, ff)", timeit(lambda: tcl.temporal_error(ff, bf, prev, cur)), px, 40)
    report(tag + "temporal_loss fwd (mask)", timeit(lambda: tcl.temporal_loss(m, cur, prev, bf)), px, 36)
    report(tag + "temporal_loss L1 fwd (mask)", timeit(lambda: tcl.temporal_loss(m, cur, prev, bf, loss="l1")), px, 36)
    report(tag + "warp_blend", timeit(lambda: tcl.warp_blend(m, prev, bf, cur)), px, 48)
    report(tag + "fused + warp/mask outputs", timeit(lambda: tcl.fused_forward(bf, prev, cur, ff=ff, want_warp=True, want_mask=True)), px, 56)
    p2 = prev.clone().requires_grad_(True)
    c2 = cur.clone().requires_grad_(True)

    def fb():
        p2.grad = None
        c2.grad = None
        tcl.temporal_loss(m, c2, p2, bf).backward()
    report(tag + "temporal_loss fwd+bwd", timeit(fb), px, 36 + 36 + 24 + 12)
    pb = prev.to(torch.bfloat16)
    cb = cur.to(torch.bfloat16)
    report(tag + "temporal_error bf16 frames", timeit(lambda: tcl.temporal_error(ff, bf, pb, cb)), px, 28)
    report(tag + "warp bf16", timeit(lambda: tcl.warp(pb, bf)), px, 20)


def run_ingest():
    N, H, W = 64, 256, 256
    block = torch.randn(N, H, W, 9, device=dev)
    report(f"ingest fc2 block[{N}] HWC9 -> planar", timeit(lambda: tcl.split_fc2_block(block)), N * H * W, 72)
    flo = torch.randn(8, 436, 1024, 2, device=dev)
    report("ingest .flo payload[8] HW2 -> planar", timeit(lambda: tcl.flow_hw2_to_planar(flo)), 8 * 436 * 1024, 16)


def run_cv2compat(N=64):
    """the generators' NumPy / OpenCV flavour on HWC device tensors (Sintel shape)"""
    cfg = tcl.synth.CONFIGS["sintel_full"]
    H, W = cfg["H"], cfg["W"]
    ff, bf = tcl.synth.make_flows(N, H, W, seed=5, max_shift=32.0, max_rot_deg=3.0, device=dev)
    img, _ = tcl.synth.make_frames(N, 3, H, W, seed=5, device=dev)
    ff, bf, img = (t.permute(0, 2, 3, 1).contiguous() for t in (ff, bf, img))
    px = N * H * W
    cc = tcl.cv2compat
    report(f"cv2compat[{N}] fb_check_flows (fused)", timeit(lambda: cc.fb_check_flows(ff, bf)), px, 20)
    report(f"cv2compat[{N}] warp_image C=3", timeit(lambda: cc.warp_image(img, bf)), px, 32)
    report(f"cv2compat[{N}] warp_flow C=2", timeit(lambda: cc.warp_flow(ff, bf)), px, 24)
    png = torch.randint(0, 256, (N, H, W), dtype=torch.uint8, device=dev)
    report(f"ingest occlusion png[{N}] u8 -> mask", timeit(lambda: tcl.sintel_occlusion_mask(png)), px, 5)


def run_upsample():
    N, H, W = 4, 55, 128                  # Sintel 436(->440)x1024 at 1/8 resolution
    flow = torch.randn(N, 2, H, W, device=dev)
    mask = torch.randn(N, 576, H, W, device=dev)
    px = N * 64 * H * W
    report(f"upsample_flow[{N}] fused kernel", timeit(lambda: tcl.upsample_flow(flow, mask)), px, 44)
    # (the op sequence it replaces -- softmax + unfold + mul + sum + permute -- measured 312 us on the same input in round 1:
    #  11x; it is test infrastructure, tests/test_gpu_parity.py::test_upsample_flow_matches_raft, and not timed from here)


def run_clip(T=257):
    cfg = tcl.synth.CONFIGS["sintel_full"]
    H, W = cfg["H"], cfg["W"]
    chunks, fr = [], []
    for s0 in range(0, T - 1, 32):
        n = min(32, T - 1 - s0)
        ff, bf = tcl.synth.make_flows(n, H, W, seed=21 + s0, max_shift=32.0, max_rot_deg=3.0, device=dev)
        chunks.append((ff, bf))
    for s0 in range(0, T, 32):
        n = min(32, T - s0)
        fr.append(tcl.synth.make_frames(n, 3, H, W, seed=21 + s0, device=dev)[0])
    ff = torch.cat([c[0] for c in chunks]); bf = torch.cat([c[1] for c in chunks]); frames = torch.cat(fr)
    prev, cur = frames[:-1].contiguous(), frames[1:].contiguous()
    px = (T - 1) * H * W
    report(f"clip[{T}] pairwise tensors (40 B/px)", timeit(lambda: tcl.temporal_error_per_pair(ff, bf, prev, cur)), px, 40)
    report(f"clip[{T}] frames stored once (28 B/px)", timeit(lambda: tcl.temporal_error_clip(frames, ff, bf)), px, 28)


if __name__ == "__main__":
    run_upsample()
    run_clip()
    run_ingest()
    run_cv2compat()
    run("sintel_full", 128)
    run("train_b16_256", 256)
    run("train_b16_256", 16)
