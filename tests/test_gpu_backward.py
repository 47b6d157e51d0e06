"""GPU suite: autograd of the drop-in warp / temporal loss against F.grid_sample's autograd
(what the reference's g_loss.backward() runs, StarGANv2AdvCon/core/solver.py:181) and the C oracle."""
import numpy as np
import pytest
import torch

from conftest import golden_files, load_npz
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu


def _inputs(tcl, B, H, W, seed, shift):
    d = torch.device("cuda:0")
    ff, bf = tcl.synth.make_flows(B, H, W, seed=seed, max_shift=shift, device=d)
    prev, cur = tcl.synth.make_frames(B, 3, H, W, seed=seed, kind="white", device=d)
    return ff, bf, prev, cur


@pytest.mark.parametrize("B,H,W,shift", [(2, 32, 48, 4.0), (16, 256, 256, 24.0), (1, 436, 1024, 32.0)])
def test_warp_backward_matches_grid_sample_autograd(tcl, B, H, W, shift):
    ff, bf, prev, cur = _inputs(tcl, B, H, W, 31, shift)
    go = torch.randn_like(prev)
    grads = []
    for fn in (tp.backward_warp, tcl.warp):
        p, f = prev.clone().requires_grad_(True), bf.clone().requires_grad_(True)
        fn(p, f).backward(go)
        grads.append((p.grad, f.grad))
    (rp, rf), (kp, kf) = grads
    # scatter-add order differs run to run on both sides: tolerance, not bits
    assert float((kp - rp).abs().max()) <= 1e-5 * max(1.0, float(rp.abs().max()))
    assert float((kf - rf).abs().max()) <= 1e-4 * max(1.0, float(rf.abs().max()))


@pytest.mark.parametrize("loss", ["l2", "l1"])
@pytest.mark.parametrize("B,H,W,shift", [(2, 32, 48, 4.0), (16, 256, 256, 24.0)])
def test_temporal_loss_backward(tcl, loss, B, H, W, shift):
    ff, bf, prev, cur = _inputs(tcl, B, H, W, 41, shift)
    mask = tcl.fbcCheckTorch(ff, bf)
    ref_fn = (lambda m, c, w: tp.tcl_l2(m, c, w)) if loss == "l2" else (lambda m, c, w: tp.tcl_l1(m, c, w))
    p, c = prev.clone().requires_grad_(True), cur.clone().requires_grad_(True)
    (ref_fn(mask, c, tp.backward_warp(p, bf)) * 100.0).backward()       # lambda_tcl = 100 (main.py:94)
    p2, c2 = prev.clone().requires_grad_(True), cur.clone().requires_grad_(True)
    val = tcl.temporal_loss(mask, c2, p2, bf, loss=loss)
    (val * 100.0).backward()
    ref_val = ref_fn(mask, cur, tp.backward_warp(prev, bf))
    assert abs(float(val) - float(ref_val)) <= 1e-5 * float(ref_val)
    assert float((c2.grad - c.grad).abs().max()) <= 1e-5 * float(c.grad.abs().max())
    assert float((p2.grad - p.grad).abs().max()) <= 1e-5 * float(p.grad.abs().max())


def test_fs_warp_backward(tcl):
    ff, bf, prev, cur = _inputs(tcl, 2, 48, 64, 51, 10.0)
    go = torch.randn_like(prev)
    p = prev.clone().requires_grad_(True)
    tp.validity_warp(p, bf).backward(go)
    p2 = prev.clone().requires_grad_(True)
    tcl.fs_warp(p2, bf).backward(go)
    assert float((p2.grad - p.grad).abs().max()) <= 1e-5 * float(p.grad.abs().max())


@pytest.mark.parametrize("path", golden_files("ref_cpu_"))
def test_backward_matches_reference_autograd_golden(tcl, path):
    g = load_npz(path)
    d = torch.device("cuda:0")
    prev = torch.from_numpy(g["prev"]).to(d).requires_grad_(True)
    cur = torch.from_numpy(g["cur"]).to(d).requires_grad_(True)
    bf = torch.from_numpy(g["bf"]).to(d)
    mask = torch.from_numpy(g["mask"]).to(d)
    tcl.temporal_loss(mask, cur, prev, bf).backward()
    assert np.allclose(cur.grad.cpu().numpy(), g["grad_cur"], rtol=1e-4, atol=1e-9)
    assert np.allclose(prev.grad.cpu().numpy(), g["grad_prev"], rtol=1e-3, atol=1e-8)
    f = bf.clone().requires_grad_(True)
    p = torch.from_numpy(g["prev"]).to(d)
    ((mask * (torch.from_numpy(g["cur"]).to(d) - tcl.warp(p, f))) ** 2).mean().backward()
    assert np.allclose(f.grad.cpu().numpy(), g["grad_flow"], rtol=1e-3, atol=1e-7)


def test_congan_soft_mask_and_scalar_masked_l1(tcl):
    """ConGAN/models/cycle_gan_model.py:136-137,298: exp(-50*|real2 - warp(real1)|.mean()) and mask*|fuse - warp|.mean()."""
    ff, bf, prev, cur = _inputs(tcl, 4, 64, 96, 61, 6.0)
    cur = prev + 0.01 * cur          # close frames: the soft mask is not vanishingly small
    ref = torch.exp(-50 * torch.abs(cur - tp.backward_warp(prev, bf)).mean())
    c = cur.clone().requires_grad_(True)
    got = tcl.generateMask(c, prev, bf)
    assert abs(float(got) - float(ref)) <= 1e-5 * float(ref)
    c_ref = cur.clone().requires_grad_(True)
    torch.exp(-50 * torch.abs(c_ref - tp.backward_warp(prev, bf)).mean()).backward()
    got.backward()
    assert float((c.grad - c_ref.grad).abs().max()) <= 1e-5 * float(c_ref.grad.abs().max())


@pytest.mark.parametrize("B,C,H,W,shift,soft", [(4, 3, 64, 96, 8.0, False), (2, 3, 256, 256, 24.0, False), (2, 32, 64, 64, 6.0, True), (1, 5, 37, 53, 20.0, True)])
def test_learning_based_temporal_loss_uses_fs_lib_warp(tcl, B, C, H, W, shift, soft):
    """fs_ruder.py:97 / fs_huang.py:55-56: ((mask*(warp(prev, flow) - cur))**2).mean() with ``from fs_lib import warp`` (the
    validity-masked warp), and ReCoNet's feature-level term fs_reconet.py:56-61 (the same on C-channel feature maps under a
    bilinearly resized, i.e. soft, mask): value within 1e-5 relative, gradients to both frames within 1e-5."""
    d = torch.device("cuda:0")
    ff, bf = tcl.synth.make_flows(B, H, W, seed=71, max_shift=shift, device=d)
    prev, cur = (t.to(d) for t in tcl.synth.make_frames(B, C, H, W, seed=71, kind="smooth"))
    mask = tcl.fbcCheckTorch(ff, bf)
    if soft:   # fs_reconet.py:59: the mask goes through a bilinear resize first
        mask = torch.nn.functional.interpolate(torch.nn.functional.avg_pool2d(mask, 2), size=(H, W), mode="bilinear")
        assert 0 < int(((mask > 0) & (mask < 1)).sum())
    p, c = prev.clone().requires_grad_(True), cur.clone().requires_grad_(True)
    want = tp.tcl_l2(mask, c, tp.validity_warp(p, bf))
    (want * 10.0).backward()
    p2, c2 = prev.clone().requires_grad_(True), cur.clone().requires_grad_(True)
    got = tcl.temporal_loss(mask, c2, p2, bf, validity=True)
    (got * 10.0).backward()
    assert abs(float(got) - float(want)) <= 1e-5 * float(want)
    assert float((c2.grad - c.grad).abs().max()) <= 1e-5 * float(c.grad.abs().max())
    assert float((p2.grad - p.grad).abs().max()) <= 1e-5 * float(p.grad.abs().max())
    # the validity factor matters on these inputs (the flow leaves the frame somewhere): not the plain warp's loss
    plain = tp.tcl_l2(mask, cur, tp.backward_warp(prev, bf))
    assert float(plain) != float(want)


# ------------------------------------------------------------------ warp chains of the learning-based trainers
@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,shift", [(2, 64, 96, 5.0), (1, 37, 53, 30.0), (4, 256, 256, 12.0)])
def test_reconet_output_temporal_loss_matches_the_reference_expression(tcl, B, H, W, shift):
    """fs_reconet.py:63-69 in one fused pass: value within 1e-5 relative, gradients to both stylised frames within 1e-5 of
    autograd through the reference's op sequence (oracle/torch_port.py on the same GPU)."""
    d = torch.device("cuda")
    ff, bf, s1, s2 = _inputs(tcl, B, H, W, 21, shift)
    i1, i2 = (t.to(d) for t in tcl.synth.make_frames(B, 3, H, W, seed=22, kind="smooth"))
    mask = tcl.fbcCheckTorch(ff, bf)
    a1, a2 = s1.clone().requires_grad_(True), s2.clone().requires_grad_(True)
    want = tp.reconet_output_loss(mask, a2, a1, i2, i1, bf)
    gw1, gw2 = torch.autograd.grad(want * 100.0, (a1, a2))
    b1, b2 = s1.clone().requires_grad_(True), s2.clone().requires_grad_(True)
    got = tcl.reconet_output_temporal_loss(mask, b2, b1, i2, i1, bf)
    assert got.dim() == 0 and abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    gg1, gg2 = torch.autograd.grad(got * 100.0, (b1, b2))
    scale = float(gw2.abs().max())
    assert float((gg2 - gw2).abs().max()) <= 1e-5 * max(scale, 1e-12)
    assert float((gg1 - gw1).abs().max()) <= 1e-5 * max(float(gw1.abs().max()), scale, 1e-12)
    # no mask = a mask of ones
    ones = torch.ones_like(mask)
    want1 = float(tp.reconet_output_loss(ones, s2, s1, i2, i1, bf))
    assert abs(float(tcl.reconet_output_temporal_loss(None, s2, s1, i2, i1, bf)) - want1) <= 1e-5 * want1
    with pytest.raises(RuntimeError):
        tcl.reconet_output_temporal_loss(mask, b2, b1, i2, i1, bf.clone().requires_grad_(True))


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,shift", [(2, 64, 96, 5.0), (1, 37, 53, 30.0)])
def test_ruder_network_input_is_warp_plus_cat(tcl, B, H, W, shift):
    """fs_ruder.py:47-50: cat((img, mask, warp(styled_prev, flow)), 1) bit for bit, the warped frame as well, and the gradient
    the chain sends back to the previous stylised frame through both uses of the warp."""
    d = torch.device("cuda")
    ff, bf, sp, img = _inputs(tcl, B, H, W, 33, shift)
    mask = tcl.fbcCheckTorch(ff, bf)
    want_cat, want_w = tp.ruder_input(img, mask, sp, bf)
    cat, warped = tcl.ruder_network_input(img, mask, sp, bf)
    assert cat.shape == (B, 7, H, W) and torch.equal(cat, want_cat) and torch.equal(warped, want_w)
    a, b = sp.clone().requires_grad_(True), sp.clone().requires_grad_(True)
    wc, ww = tp.ruder_input(img, mask, a, bf)
    gc, gw = tcl.ruder_network_input(img, mask, b, bf)
    w1 = torch.randn_like(wc)
    ga = torch.autograd.grad((wc * w1).sum() + 0.5 * ((mask * (ww - img)) ** 2).mean(), a)[0]
    gb = torch.autograd.grad((gc * w1).sum() + 0.5 * ((mask * (gw - img)) ** 2).mean(), b)[0]
    assert float((ga - gb).abs().max()) <= 1e-5 * max(float(ga.abs().max()), 1e-12)
