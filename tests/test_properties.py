"""Property tests the reference implies: an analytic frame moved by an affine map must come back
under the warp, a consistent flow pair must pass the forward-backward check everywhere but on the image border.
The same checks run on the oracle (CPU suite) and on the kernels (GPU suite).

The reference's own sanity check is of this kind: coco-generation.py:211-227,272-300 warps an affinely transformed frame
back and accumulates the masked MSE against the ground truth."""
import math

import numpy as np
import pytest
import torch


def analytic_frame(xs, ys):
    """smooth, band-limited test image f(x, y) with three channels in [-1, 1]"""
    return np.stack([np.sin(xs / 23.0) * np.cos(ys / 17.0), np.sin((xs + 2 * ys) / 31.0), np.cos(xs / 41.0 - ys / 13.0)]).astype(np.float32)


def affine_case(tcl, B, H, W, seed):
    ff, bf = tcl.synth.make_flows(B, H, W, seed=seed, max_shift=6.0, max_rot_deg=3.0, n_rects=0, residual=0.0)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    prev = np.stack([analytic_frame(xs, ys)] * B)
    # where warp() samples: the reference normalises by size-1 and samples with align_corners=False (flowtools.py:28-32)
    ix = (xs[None] + bf[:, 0].numpy()) * (W / (W - 1.0)) - 0.5
    iy = (ys[None] + bf[:, 1].numpy()) * (H / (H - 1.0)) - 0.5
    want = np.stack([analytic_frame(ix[b], iy[b]) for b in range(B)])
    inside = (ix >= 0) & (ix <= W - 1) & (iy >= 0) & (iy <= H - 1)
    return ff, bf, torch.from_numpy(prev), want, inside


def check_warp_recovers_the_analytic_frame(warped, want, inside):
    err = np.abs(warped - want)[np.broadcast_to(inside[:, None], want.shape)]
    # bilinear interpolation of a band-limited image: second-order error, h = 1 px, |f''| <= ~1/13^2 * few
    assert err.max() < 5e-3, err.max()
    assert err.mean() < 1e-3


def check_consistent_flows_pass(mask, H, W):
    m = mask[:, 0]
    assert set(np.unique(m).tolist()) <= {0.0, 1.0}
    # zero-padded gradient: any non-zero flow on the border trips the motion-boundary test (flowtools.py:13-14,53)
    assert m[:, 0, :].max() == 0 and m[:, -1, :].max() == 0 and m[:, :, 0].max() == 0 and m[:, :, -1].max() == 0
    # the interior of an exactly consistent affine pair passes except where the backward flow points out of the frame
    assert m[:, 8:-8, 8:-8].mean() > 0.97


def test_oracle_affine_properties(tcl, oracle_mod):
    B, H, W = 2, 96, 160
    ff, bf, prev, want, inside = affine_case(tcl, B, H, W, seed=11)
    for variant in (oracle_mod.ATEN_CPU, oracle_mod.ATEN_CUDA):
        check_warp_recovers_the_analytic_frame(oracle_mod.warp(prev.numpy(), bf.numpy(), variant), want, inside)
        check_consistent_flows_pass(oracle_mod.fbcheck(ff.numpy(), bf.numpy(), variant=variant), H, W)
    # a frame pair that IS the affine motion has (nearly) zero temporal error inside the mask; an unrelated frame does not
    sums = oracle_mod.temporal_error_sums(ff.numpy(), bf.numpy(), prev.numpy(), want.astype(np.float32))
    unrelated = oracle_mod.temporal_error_sums(ff.numpy(), bf.numpy(), prev.numpy(), np.ascontiguousarray(want[:, :, ::-1, ::-1]))
    assert (np.sqrt(sums / (3 * H * W)) < 1e-2).all() and (np.sqrt(unrelated / (3 * H * W)) > 0.1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W", [(2, 96, 160), (1, 436, 1024)])
def test_kernels_affine_properties(tcl, B, H, W):
    d = torch.device("cuda:0")
    ff, bf, prev, want, inside = affine_case(tcl, B, H, W, seed=11)
    cur = torch.from_numpy(want.astype(np.float32))
    r = tcl.fused_forward(bf.to(d), prev.to(d), cur.to(d), ff=ff.to(d), want_warp=True, want_mask=True)
    check_warp_recovers_the_analytic_frame(r.warp.cpu().numpy(), want, inside)
    check_consistent_flows_pass(r.mask.cpu().numpy(), H, W)
    assert float(r.pair_vals.max()) < 1e-2                                     # the pair IS the motion: no temporal error
    flipped = torch.flip(cur, dims=(2, 3)).contiguous().to(d)
    assert float(tcl.temporal_error_per_pair(ff.to(d), bf.to(d), prev.to(d), flipped).min()) > 0.1
    # linearity of the masked L2 sums in the squared frame scale: scaling both frames by 2 scales the RMSE by 2 (exactly, powers of two)
    r2 = tcl.fused_forward(bf.to(d), (2 * prev).to(d), (2 * cur).to(d), ff=ff.to(d))
    assert torch.equal(r2.pair_sums, 4 * tcl.fused_forward(bf.to(d), prev.to(d), cur.to(d), ff=ff.to(d)).pair_sums)
