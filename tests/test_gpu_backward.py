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
