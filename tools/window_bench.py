"""BASELINE config 4 (1080p, bf16 frames, 4-frame temporal window, both directions) as independent pairs vs window mode
(frames and flow fields stored once, interleaved tiles).  Kernel-only timing, tuning aid."""
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


def sustained(fn, secs=2.0):
    """Gpix-agnostic: ms per call after `secs` of back-to-back calls (power-capped clock)."""
    import time
    t0 = time.time()
    while time.time() - t0 < secs:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    return timeit(fn, n=40, warm=0)


def run(name, T, window, dtype):
    cfg = tcl.synth.CONFIGS[name]
    H, W = cfg["H"], cfg["W"]
    idx = tcl.window_evaluations(T, window)
    J, E = idx["field_t"].numel(), idx["prev_index"].numel()
    ffs, bfs = [], []
    for s in range(0, J, 8):
        n = min(8, J - s)
        f, b = tcl.synth.make_flows(n, H, W, seed=500 + s, max_shift=cfg["max_shift"], max_rot_deg=cfg["max_rot_deg"], device=dev)
        ffs.append(f); bfs.append(b)
    ff, bf = torch.cat(ffs), torch.cat(bfs)
    bank = torch.stack([bf, ff], dim=1).reshape(2 * J, 2, H, W).contiguous()
    del ffs, bfs, ff, bf
    frames, _ = tcl.synth.make_frames(T, 3, H, W, seed=501, device=dev, dtype=dtype)
    li = lambda k: idx[k].long().to(dev)
    ii = lambda k: idx[k].to(dev)
    # independent pairs: every evaluation owns its two flows and two frames (28 B/px bf16, 40 fp32)
    ffm, bfm = bank[li("ff_index")].contiguous(), bank[li("bf_index")].contiguous()
    pm, cm = frames[li("prev_index")].contiguous(), frames[li("cur_index")].contiguous()
    esz = frames.element_size()
    px = E * H * W
    f_ind = lambda: tcl.fused_forward(bfm, pm, cm, ff=ffm)
    t_ind = timeit(f_ind)
    s_ind = sustained(f_ind)
    del ffm, bfm, pm, cm, f_ind
    t_idx = timeit(lambda: tcl.fused_forward(bank, frames, frames, ff=bank, prev_index=ii("prev_index"), cur_index=ii("cur_index"),
                                            bf_index=ii("bf_index"), ff_index=ii("ff_index"), validate_index=False))
    f_win = lambda: tcl.fused_forward(bank, frames, frames, ff=bank, prev_index=ii("prev_index"), cur_index=ii("cur_index"),
                                      bf_index=ii("bf_index"), ff_index=ii("ff_index"), validate_index=False, pair_group=idx["group"])
    t_win = timeit(f_win)
    s_win = sustained(f_win)
    bpp_ind = 16 + 6 * esz
    stored = (2 * J * 8 + T * 3 * esz) * H * W
    print(f"{name} {W}x{H} {str(dtype)[6:]} T={T} window={window}: {E} evaluations, {2 * J} flow fields, {T} frames", flush=True)
    print(f"  independent pairs ({bpp_ind} B/px):           {t_ind * 1e3:9.1f} us  {px / t_ind / 1e6:7.1f} Gpix/s")
    print(f"  window mode, stored once ({stored / px:.1f} B/px): {t_idx * 1e3:9.1f} us  {px / t_idx / 1e6:7.1f} Gpix/s  (pair-major tiles)")
    print(f"  window mode + interleaved tiles:          {t_win * 1e3:9.1f} us  {px / t_win / 1e6:7.1f} Gpix/s")
    print(f"  sustained (2 s of load first): independent {px / s_ind / 1e6:7.1f} Gpix/s, window mode {px / s_win / 1e6:7.1f} Gpix/s")


if __name__ == "__main__":
    run("hd1080_window", 10, 4, torch.bfloat16)
    run("hd1080_window", 10, 4, torch.float32)
    run("sintel_full", 20, 4, torch.float32)
