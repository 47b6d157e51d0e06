"""TEST INFRASTRUCTURE ONLY -- the parity oracle for the temporal-consistency hot path.

``oracle/tcl_oracle.c`` is a plain-C CPU restatement of the reference's algorithm and
``oracle/torch_port.py`` an op-for-op ATen restatement; see oracle/README.md.  Nothing under
``gan-based-video-style-transfer_b200/`` imports this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "tcl_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB = os.path.join(_OUT_DIR, "libtcl_oracle.so")

# variant bits, mirrored from tcl_oracle.c
V_NORM_RECIP, V_UNNORM_CUDA, V_UNNORM_FMA, V_WEIGHT_CUDA, V_ACC_FMA, V_SQ_FMA = (1 << i for i in range(6))
ATEN_CUDA = V_NORM_RECIP | V_UNNORM_CUDA | V_UNNORM_FMA | V_WEIGHT_CUDA | V_ACC_FMA
ATEN_CPU = V_UNNORM_FMA | V_ACC_FMA  # bit-exact vs torch 2.11 CPU, see tests/test_oracle_pinning.py
FLAG_OCC, FLAG_MOB = 1, 2


def build(force=False):
    """gcc the C restatement into oracle/_build/ (no reference sources involved)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
           "-fopenmp", "-o", _LIB, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def central_diff(x):
    x, px = _f32(x)
    B, H, W = x.shape
    out = np.empty((2, B, H, W), np.float32)
    lib().oracle_central_diff(px, B, H, W, out.ctypes.data_as(ctypes.c_void_p))
    return out


def sample_coords(f, variant=ATEN_CUDA):
    f, pf = _f32(f)
    B, _, H, W = f.shape
    ix = np.empty((B, H, W), np.float32)
    iy = np.empty((B, H, W), np.float32)
    lib().oracle_sample_coords(pf, B, H, W, ix.ctypes.data_as(ctypes.c_void_p),
                               iy.ctypes.data_as(ctypes.c_void_p), variant)
    return ix, iy


def warp(x, f, variant=ATEN_CUDA):
    x, px = _f32(x)
    f, pf = _f32(f)
    B, C, H, W = x.shape
    out = np.empty_like(x)
    lib().oracle_warp(px, pf, B, C, H, W, out.ctypes.data_as(ctypes.c_void_p), variant)
    return out


def validity_warp(x, f, variant=ATEN_CUDA):
    x, px = _f32(x)
    f, pf = _f32(f)
    B, C, H, W = x.shape
    out = np.empty_like(x)
    lib().oracle_validity_warp(px, pf, B, C, H, W, out.ctypes.data_as(ctypes.c_void_p), variant)
    return out


def fbcheck(ff, bf, flags=FLAG_OCC | FLAG_MOB, variant=ATEN_CUDA, margins=False):
    ff, pff = _f32(ff)
    bf, pbf = _f32(bf)
    B, _, H, W = bf.shape
    mask = np.empty((B, 1, H, W), np.float32)
    if margins:
        mo = np.empty((B, H, W), np.float32)
        mm = np.empty((B, H, W), np.float32)
        lib().oracle_fbcheck(pff, pbf, B, H, W, flags, mask.ctypes.data_as(ctypes.c_void_p),
                             mo.ctypes.data_as(ctypes.c_void_p), mm.ctypes.data_as(ctypes.c_void_p), variant)
        return mask, mo, mm
    lib().oracle_fbcheck(pff, pbf, B, H, W, flags, mask.ctypes.data_as(ctypes.c_void_p), None, None, variant)
    return mask


def masked_sums(mask, cur, warped, mode=0):
    mask, pm = _f32(mask)
    cur, pc = _f32(cur)
    warped, pw = _f32(warped)
    B, C, H, W = cur.shape
    sums = np.empty((B,), np.float64)
    lib().oracle_masked_sums(pm, pc, pw, B, C, H, W, mode, sums.ctypes.data_as(ctypes.c_void_p))
    return sums


def blend(mask, warped, img):
    mask, pm = _f32(mask)
    warped, pw = _f32(warped)
    img, pi = _f32(img)
    B, C, H, W = img.shape
    out = np.empty_like(img)
    lib().oracle_blend(pm, pw, pi, B, C, H, W, out.ctypes.data_as(ctypes.c_void_p))
    return out


def temporal_error_sums(ff, bf, prev, cur, flags=FLAG_OCC | FLAG_MOB, variant=ATEN_CUDA):
    ff, pff = _f32(ff)
    bf, pbf = _f32(bf)
    prev, pp = _f32(prev)
    cur, pc = _f32(cur)
    B, C, H, W = cur.shape
    sums = np.empty((B,), np.float64)
    lib().oracle_temporal_error_sums(pff, pbf, pp, pc, B, C, H, W, flags,
                                     sums.ctypes.data_as(ctypes.c_void_p), variant)
    return sums


def warp_bwd(grad_out, x, f, variant=ATEN_CUDA, need_grad_x=True, need_grad_f=True):
    grad_out, pg = _f32(grad_out)
    x, px = _f32(x)
    f, pf = _f32(f)
    B, C, H, W = x.shape
    gx = np.empty_like(x) if need_grad_x else None
    gf = np.empty_like(f) if need_grad_f else None
    lib().oracle_warp_bwd(pg, px, pf, B, C, H, W,
                          gx.ctypes.data_as(ctypes.c_void_p) if need_grad_x else None,
                          gf.ctypes.data_as(ctypes.c_void_p) if need_grad_f else None, variant)
    return gx, gf
