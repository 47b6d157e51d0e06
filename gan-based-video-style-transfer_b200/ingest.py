"""Dataset ingest for the temporal-consistency path: interleaved blocks -> the planar tensors the kernels consume.

Upstream formats (paths relative to the reference repository):

* the 9-channel HWC ``.npy`` blocks ``[img1 3 | img2 3 | mask 1 | flow 2]`` of the FlyingChairs2 / Hollywood2
  training sets (``methods/GAN-based/StarGANv2AdvCon/core/data_loader.py:243-245``,
  ``methods/learning-based/datasets.py:52-54``): the reference slices them with ``np.moveaxis(np_data[:,:,6:7], 2, 0)``
  per sample on the host;
* ``.flo`` files (``utils/flowlib.py:33-48``): magic ``PIEH``, int32 width, int32 height, H x W x 2 float32;
* Sintel's ground-truth occlusion PNGs (``utils/sintel_dataset.py:64-65``): ``mask = 1.0 - imread(png)/255.0``;
* the long-term ``.npy`` blocks ``[flow 2 | mask 1]`` HWC (``utils/sintel_dataset.py:76-83``): ``hwc_split(block, LT_LAYOUT)``.

Here the raw interleaved block is uploaded once and split on the GPU by ``tclb200_hwc_split`` (one pass, coalesced on both
sides).  File reading itself stays plain numpy: it is I/O, not arithmetic.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from ._cabi import check

FC2_LAYOUT = (("img1", 0, 3), ("img2", 3, 3), ("mask", 6, 1), ("flow", 7, 2))
LT_LAYOUT = (("flow", 0, 2), ("mask", 2, 1))   # utils/sintel_dataset.py:76-78: data[0,:,:,:2], data[0,:,:,2]


def hwc_split(block, parts):
    """``block`` (N,H,W,Cs) or (H,W,Cs) float32 CUDA tensor -> dict name -> (N,Cd,H,W) planar tensors.

    ``parts`` is a sequence of ``(name, first_channel, n_channels)``."""
    if not block.is_cuda:
        raise RuntimeError("tcl_b200: hwc_split expects a CUDA tensor (upload the raw block, it is split on the GPU)")
    if block.dim() == 3:
        block = block.unsqueeze(0)
    if block.dim() != 4 or block.dtype != torch.float32:
        raise RuntimeError(f"tcl_b200: hwc_split expects float32 (N,H,W,C), got {tuple(block.shape)} {block.dtype}")
    block = block.contiguous()
    N, H, W, Cs = block.shape
    parts = list(parts)
    if not 1 <= len(parts) <= 8:
        raise RuntimeError("tcl_b200: 1..8 outputs per call")
    outs = {name: torch.empty((N, cd, H, W), dtype=torch.float32, device=block.device) for name, _, cd in parts}
    n = len(parts)
    dst = (ctypes.c_void_p * n)(*[outs[name].data_ptr() for name, _, _ in parts])
    c0 = (ctypes.c_int * n)(*[int(a) for _, a, _ in parts])
    cd = (ctypes.c_int * n)(*[int(b) for _, _, b in parts])
    with torch.cuda.device(block.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(_cabi.lib().tclb200_hwc_split(ctypes.c_void_p(block.data_ptr()), N, H, W, Cs, n, dst, c0, cd, stream))
    return outs


def split_fc2_block(block):
    """The reference's 9-channel training block -> ``img1, img2 (N,3,H,W), mask (N,1,H,W), flow (N,2,H,W)``."""
    o = hwc_split(block, FC2_LAYOUT)
    return o["img1"], o["img2"], o["mask"], o["flow"]


def flow_hw2_to_planar(flow_hw2):
    """(N,H,W,2) or (H,W,2) interleaved flow (a .flo payload) -> (N,2,H,W)."""
    return hwc_split(flow_hw2, (("flow", 0, 2),))["flow"]


def read_flo(path):
    """``flowlib.readFlow`` for .flo files: returns the (H,W,2) float32 numpy payload (host I/O only)."""
    with open(path, "rb") as f:
        if f.read(4) != b"PIEH":
            raise Exception("Flow file header does not contain PIEH")
        w = int(np.fromfile(f, np.int32, 1).squeeze())
        h = int(np.fromfile(f, np.int32, 1).squeeze())
        return np.fromfile(f, np.float32, w * h * 2).reshape((h, w, 2)).astype(np.float32)


def write_flo(path, flow_hw2):
    """``flowlib.writeFlow``: (H,W,2) float32 -> .flo."""
    flow_hw2 = np.asarray(flow_hw2, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(b"PIEH")
        np.array([flow_hw2.shape[1], flow_hw2.shape[0]], dtype=np.int32).tofile(f)
        flow_hw2.tofile(f)


def load_flo_planar(path, device="cuda"):
    """.flo file -> (1,2,H,W) CUDA tensor (upload interleaved, de-interleave on the GPU)."""
    return flow_hw2_to_planar(torch.from_numpy(read_flo(path)).to(device))


def sintel_occlusion_mask(png_u8):
    """Sintel occlusion PNG (uint8 CUDA tensor (H,W) or (N,H,W), 255 = occluded) -> mask (N,1,H,W) float32 =
    ``(1.0 - png/255.0).float()`` exactly as utils/sintel_dataset.py:64-65 computes it (float64, then ``.float()``)."""
    if not png_u8.is_cuda or png_u8.dtype != torch.uint8:
        raise RuntimeError("tcl_b200: sintel_occlusion_mask expects a uint8 CUDA tensor (upload the decoded PNG as it is)")
    if png_u8.dim() == 2:
        png_u8 = png_u8.unsqueeze(0)
    if png_u8.dim() != 3:
        raise RuntimeError(f"tcl_b200: expected (H,W) or (N,H,W), got {tuple(png_u8.shape)}")
    src = png_u8.contiguous()
    N, H, W = src.shape
    out = torch.empty((N, 1, H, W), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(_cabi.lib().tclb200_occlusion_u8_to_mask(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(out.data_ptr()), src.numel(), stream))
    return out
