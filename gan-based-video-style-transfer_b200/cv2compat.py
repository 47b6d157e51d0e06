"""NumPy / OpenCV flavour of the path, on the GPU: the functions the reference's dataset generators define.

Upstream: methods/learning-based/dataset-generation/coco-generation.py:66-113 (``warp_image``, ``warp_flow``,
``fb_check``; same in hollywood2-generation.py:63-111, sintel-generation.py:89-130, fast_style_transfer.py:824-831).
They are NOT numerically equal to ``flowtools.warp`` / ``fbcCheckTorch``: ``cv2.remap`` samples at exactly
(x+u, y+v) in 1/32-pixel fixed point, ``np.gradient`` is one-sided at the borders.  The kernels in
``csrc/tcl_cv2.cu`` reproduce those numerics bit for bit (tests/test_cv2_compat.py), on the reference's HWC layout.

Each function accepts what the reference passes -- NumPy arrays (H,W,C) -- and returns NumPy, or CUDA tensors
(H,W,C) / (N,H,W,C), which stay on the device.  There is no CPU path: NumPy inputs are copied to the current CUDA
device and back.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from ._cabi import MOB, OCC, check


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_device(a, name):
    """-> (tensor (N,H,W,C) fp32 contiguous on CUDA, was_numpy, had_batch_dim)"""
    was_numpy = isinstance(a, np.ndarray)
    if was_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError("tcl_b200.cv2compat: no CUDA device; this path has no CPU implementation")
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    elif torch.is_tensor(a):
        if not a.is_cuda:
            raise RuntimeError(f"tcl_b200.cv2compat: {name} must be a NumPy array or a CUDA tensor")
        t = a.float().contiguous()
    else:
        raise TypeError(f"tcl_b200.cv2compat: {name} must be a NumPy array or a CUDA tensor")
    if t.dim() == 2:
        t = t.unsqueeze(-1)
    batched = t.dim() == 4
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise RuntimeError(f"tcl_b200.cv2compat: {name} must be (H,W,C) or (N,H,W,C), got {tuple(a.shape)}")
    return t, was_numpy, batched


def _flow(f, name):
    t, was_numpy, batched = _to_device(f, name)
    if t.shape[-1] != 2:
        raise RuntimeError(f"tcl_b200.cv2compat: {name} must end in 2 channels (u, v), got {tuple(f.shape)}")
    return t, was_numpy, batched


def warp_flow(A, flow):
    """``cv2.remap(A, x + flow[...,0], y + flow[...,1], cv2.INTER_LINEAR)`` (coco-generation.py:86-94)."""
    img, np_in, batched = _to_device(A, "A")
    fl, _, _ = _flow(flow, "flow")
    if img.shape[:3] != fl.shape[:3]:
        # the reference's assert (coco-generation.py:88)
        raise AssertionError("dimension error: input and flow size do not match")
    N, H, W, C = img.shape
    out = torch.empty_like(img)
    with torch.cuda.device(img.device):
        check(_cabi.lib().tclb200_cv2_remap(ctypes.c_void_p(img.data_ptr()), ctypes.c_void_p(fl.data_ptr()),
                                            ctypes.c_void_p(out.data_ptr()), N, H, W, C, _stream()))
    if not batched:
        out = out[0]
    if np_in:
        return out.cpu().numpy().reshape(A.shape)
    return out.reshape(A.shape) if A.dim() == 2 else out


def warp_image(A, flow):
    """coco-generation.py:66-84: the same remap on an image, result reshaped like ``A``."""
    return warp_flow(A, flow)


def _fb_check(ff, bf, motion_boundaries, prewarped, return_near=False):
    f, np_in, batched = _flow(ff, "w_warp" if prewarped else "ff")
    b, _, _ = _flow(bf, "w_back" if prewarped else "bf")
    if f.shape != b.shape:
        raise RuntimeError(f"tcl_b200.cv2compat: flows {tuple(f.shape)} and {tuple(b.shape)} do not match")
    N, H, W, _ = b.shape
    mask = torch.empty((N, H, W), dtype=torch.float32, device=b.device)
    near = torch.zeros(1, dtype=torch.int64, device=b.device) if return_near else None
    flags = OCC | (MOB if motion_boundaries else 0)
    with torch.cuda.device(b.device):
        check(_cabi.lib().tclb200_cv2_fb_check(ctypes.c_void_p(f.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                               ctypes.c_void_p(mask.data_ptr()), N, H, W, flags, 1 if prewarped else 0,
                                               ctypes.c_void_p(near.data_ptr()) if near is not None else None, _stream()))
    if not batched:
        mask = mask[0]
    if np_in:
        mask = mask.cpu().numpy().astype(np.float64)    # the reference's weights are np.ones(...) = float64
    return (mask, near) if return_near else mask


def fb_check(w_warp, w_back, motion_boundaries=True):
    """``fb_check(w_warp, w_back)`` of the generators: weights (H,W) in {0,1}; ``w_warp`` is the already warped flow.
    ``motion_boundaries=False`` is the COCO copy, which comments the boundary test out (coco-generation.py:111)."""
    return _fb_check(w_warp, w_back, motion_boundaries, prewarped=True)


def fb_check_flows(ff, bf, motion_boundaries=True, return_near=False):
    """``fb_check(warp_flow(ff, bf), bf)`` (coco-generation.py:273-274) in one pass: the warped flow never exists."""
    return _fb_check(ff, bf, motion_boundaries, prewarped=False, return_near=return_near)
