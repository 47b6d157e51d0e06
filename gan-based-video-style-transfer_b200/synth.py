"""Seeded synthetic frame pairs + forward/backward flows for tests and bench.py.

There is no network on the build/GPU boxes, so MPI-Sintel / FlyingChairs2 and RAFT
checkpoints are unavailable; SURVEY.md section 8(d) defines the synthetic stand-ins.  The flow
model follows the reference's own synthetic generator (a random global affine map,
``methods/learning-based/dataset-generation/coco-generation.py:151-173,211-224``) and adds
independently moving rectangles (interior motion boundaries + true occlusions) and a smooth
low-amplitude residual, so that the occlusion mask keeps roughly 60-95 % of the pixels.

Conventions: ``bf`` is the flow t -> t-1 sampled on frame t's grid (the
flow ``warp`` consumes), ``ff`` the flow t-1 -> t on frame t-1's grid; channel 0 = u (x), 1 = v (y).
Nothing here is on the measured path.
"""
import math

import torch
import torch.nn.functional as F

# Sintel training sequences' frame counts (23 sequences, 1064 frames -> 1041 short-term pairs)
SINTEL_TRAIN_FRAMES = [50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 21, 50, 50, 50, 50, 50,
                       50, 50, 33, 10]
assert sum(SINTEL_TRAIN_FRAMES) == 1064 and len(SINTEL_TRAIN_FRAMES) == 23


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _lowfreq(shape, H, W, cells, g, device):
    """Band-limited noise in U[-1,1]: bicubic upsampling of a coarse random grid."""
    ch, cw = max(2, H // cells + 2), max(2, W // cells + 2)
    coarse = torch.rand(*shape, ch, cw, generator=g, device=device) * 2 - 1
    lead = coarse.shape[:-2]
    out = F.interpolate(coarse.reshape(-1, 1, ch, cw), size=(H, W), mode="bicubic", align_corners=True)
    return out.reshape(*lead, H, W)


def make_flows(B, H, W, seed=1234, max_shift=32.0, max_rot_deg=4.0, max_scale_px=16.0, n_rects=8,
               rect_shift=12.0, residual=0.3, device="cpu"):
    """Return (ff, bf), each (B,2,H,W) fp32, approximately forward-backward consistent."""
    g = _gen(seed, device)
    r = lambda *s: torch.rand(*s, generator=g, device=device) * 2 - 1
    ys, xs = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32),
                            torch.arange(W, device=device, dtype=torch.float32), indexing="ij")
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    ang = r(B) * math.radians(max_rot_deg)
    scl = 1.0 + r(B) * (max_scale_px / max(H, W))
    sh = r(B, 2) * max_shift
    ca, sa = (scl * torch.cos(ang)).view(B, 1, 1), (scl * torch.sin(ang)).view(B, 1, 1)
    # layer k maps frame-t coords p to frame-(t-1) coords  T_k(p) = A (p - c) + c + shift + d_k
    K = n_rects
    d = torch.cat([torch.zeros(B, 1, 2, device=device), r(B, K, 2) * rect_shift], 1)  # (B,K+1,2)
    # rectangles in frame t: centre, half extents
    rc = torch.stack([torch.rand(B, K, generator=g, device=device) * W,
                      torch.rand(B, K, generator=g, device=device) * H], -1)
    rh = torch.stack([(0.04 + 0.12 * torch.rand(B, K, generator=g, device=device)) * W,
                      (0.04 + 0.12 * torch.rand(B, K, generator=g, device=device)) * H], -1)

    def fwd_map(px, py, k):  # T_k
        qx = ca * (px - cx) - sa * (py - cy) + cx + (sh[:, 0] + d[:, k, 0]).view(B, 1, 1)
        qy = sa * (px - cx) + ca * (py - cy) + cy + (sh[:, 1] + d[:, k, 1]).view(B, 1, 1)
        return qx, qy

    det = (ca * ca + sa * sa)

    def inv_map(qx, qy, k):  # T_k^{-1}
        tx = qx - cx - (sh[:, 0] + d[:, k, 0]).view(B, 1, 1)
        ty = qy - cy - (sh[:, 1] + d[:, k, 1]).view(B, 1, 1)
        px = (ca * tx + sa * ty) / det + cx
        py = (-sa * tx + ca * ty) / det + cy
        return px, py

    def in_rect(px, py, k):  # k >= 1
        c, h = rc[:, k - 1], rh[:, k - 1]
        return ((px - c[:, 0].view(B, 1, 1)).abs() <= h[:, 0].view(B, 1, 1)) & \
               ((py - c[:, 1].view(B, 1, 1)).abs() <= h[:, 1].view(B, 1, 1))

    X, Y = xs.expand(B, H, W), ys.expand(B, H, W)
    qx, qy = fwd_map(X, Y, 0)
    bu, bv = qx - X, qy - Y
    px, py = inv_map(X, Y, 0)
    fu, fv = px - X, py - Y
    for k in range(1, K + 1):
        sel = in_rect(X, Y, k)
        qx, qy = fwd_map(X, Y, k)
        bu, bv = torch.where(sel, qx - X, bu), torch.where(sel, qy - Y, bv)
        px, py = inv_map(X, Y, k)
        sel = in_rect(px, py, k)
        fu, fv = torch.where(sel, px - X, fu), torch.where(sel, py - Y, fv)
    bf = torch.stack([bu, bv], 1)
    ff = torch.stack([fu, fv], 1)
    if residual > 0:
        # amplitude itself varies smoothly over the frame so threshold crossings occur
        amp = residual * (1.0 + _lowfreq((B, 1), H, W, 96, g, device))
        bf = bf + amp * _lowfreq((B, 2), H, W, 48, g, device)
        ff = ff + amp * _lowfreq((B, 2), H, W, 48, g, device)
    return ff.contiguous(), bf.contiguous()


def make_frames(B, C, H, W, seed=1234, kind="smooth", device="cpu", dtype=torch.float32):
    """Return (prev, cur) in [-1,1] (the GAN methods' Normalize(0.5,0.5) range)."""
    g = _gen(seed + 7919, device)
    if kind == "white":
        prev = torch.rand(B, C, H, W, generator=g, device=device) * 2 - 1
        cur = torch.rand(B, C, H, W, generator=g, device=device) * 2 - 1
    elif kind == "smooth":
        prev = _lowfreq((B, C), H, W, 8, g, device).clamp_(-1, 1)
        cur = (0.9 * prev.roll(shifts=(1, 2), dims=(2, 3)) + 0.1 * _lowfreq((B, C), H, W, 8, g, device)).clamp_(-1, 1)
    else:
        raise ValueError("kind must be 'smooth' or 'white'")
    return prev.to(dtype).contiguous(), cur.to(dtype).contiguous()


# displacement scale per BASELINE.json config
CONFIGS = {
    "sintel_clip":   dict(H=436, W=1024, C=3, pairs=49, max_shift=32.0, max_rot_deg=3.0, dtype="fp32"),
    "train_b16_256": dict(H=256, W=256, C=3, pairs=16, max_shift=24.0, max_rot_deg=6.0, dtype="fp32"),
    "sintel_full":   dict(H=436, W=1024, C=3, pairs=1041, max_shift=32.0, max_rot_deg=3.0, dtype="fp32"),
    "hd1080_window": dict(H=1080, W=1920, C=3, pairs=6, max_shift=56.0, max_rot_deg=2.0, dtype="bf16"),
    "uhd4k_stress":  dict(H=2160, W=3840, C=3, pairs=1, max_shift=224.0, max_rot_deg=2.0, dtype="fp32"),
}


def make_config_batch(name, n_pairs=None, seed=None, device="cpu", frame_kind="smooth"):
    """Synthetic batch for a named BASELINE.json config: dict(ff,bf,prev,cur)."""
    cfg = CONFIGS[name]
    idx = list(CONFIGS).index(name)
    n = cfg["pairs"] if n_pairs is None else n_pairs
    seed = 1234 + 1000 * idx if seed is None else seed
    ff, bf = make_flows(n, cfg["H"], cfg["W"], seed=seed, max_shift=cfg["max_shift"],
                        max_rot_deg=cfg["max_rot_deg"], device=device)
    dt = torch.bfloat16 if cfg["dtype"] == "bf16" else torch.float32
    prev, cur = make_frames(n, cfg["C"], cfg["H"], cfg["W"], seed=seed, kind=frame_kind, device=device, dtype=dt)
    return dict(ff=ff, bf=bf, prev=prev, cur=cur)
