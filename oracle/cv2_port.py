"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's dataset-generation flavour of the path.

The dataset generators restate warp / forward-backward check with NumPy + OpenCV
(methods/learning-based/dataset-generation/coco-generation.py:66-113, hollywood2-generation.py:63-111,
sintel-generation.py:89-130): ``warp_flow`` / ``warp_image`` = ``cv2.remap(A, x+u, y+v, cv2.INTER_LINEAR)`` and
``fb_check`` = the thresholds of ``fbcCheckTorch`` on ``np.linalg.norm`` / ``np.gradient``.  Numerically this is NOT
the torch path: the sample position is exactly (x+u, y+v), the gradient is one-sided at the
borders, and the arithmetic lives in a third-party dependency that is not vendored:

* **OpenCV** (``cv2``; the reference pins no version, this container has 4.13.0).  ``cv2.remap`` with float32 maps and
  ``INTER_LINEAR`` (modules/imgproc/src/imgwarp.cpp, ``remap`` -> ``remapBilinear<Cast<float,float>, RemapNoVec, float>``):
  coordinates are converted to fixed point with 5 fractional bits, ``s = cvRound(coord * 32)`` (round half to even),
  integer part ``s >> 5``, fraction index ``s & 31``; the four weights come from a 32x32 table
  ``(1 - fy)(1 - fx), (1 - fy) fx, fy (1 - fx), fy fx`` with ``f = index / 32`` (all exact in fp32); the value is
  ``v00*w0 + v01*w1 + v10*w2 + v11*w3`` evaluated left to right in fp32 without contraction; taps outside the image
  contribute the border constant 0 (``BORDER_CONSTANT``, the default).
* **NumPy** ``np.linalg.norm(x, axis)`` = ``sqrt(add.reduce(x*x))`` in the array's dtype (fp32), ``**2.0`` squares it
  again; ``np.gradient`` = central differences halved in the interior, one-sided differences at the two borders.

Pinned by tests/test_cv2_compat.py (CPU suite): ``remap_linear`` against ``cv2.remap`` itself, bit for bit, and
``fb_check`` / ``warp_flow`` / ``warp_image`` against the reference's own functions executed from their source
(the generator scripts import ``imageio``, which is absent, so the functions are loaded by ``exec`` of the def blocks).
Only tests/ may import this module.
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS


def remap_linear(A, x, y):
    """``cv2.remap(A, x, y, cv2.INTER_LINEAR)`` for float32 ``A`` (H,W) or (H,W,C) and float32 maps (H,W)."""
    A = np.asarray(A, np.float32)
    squeeze = A.ndim == 2
    if squeeze:
        A = A[..., None]
    H, W, _ = A.shape
    sx = np.rint(np.asarray(x, np.float32) * np.float32(INTER_TAB_SIZE))    # cvRound: half to even
    sy = np.rint(np.asarray(y, np.float32) * np.float32(INTER_TAB_SIZE))
    # far outside (or non-finite) either way: no tap can land in an image whose sides are below 32768
    bad = ~(np.abs(sx) < 2.0 ** 30) | ~(np.abs(sy) < 2.0 ** 30)
    sx = np.where(bad, 0, sx).astype(np.int64)
    sy = np.where(bad, 0, sy).astype(np.int64)
    x0, y0 = sx >> INTER_BITS, sy >> INTER_BITS
    fx = (sx & (INTER_TAB_SIZE - 1)).astype(np.float32) / np.float32(INTER_TAB_SIZE)
    fy = (sy & (INTER_TAB_SIZE - 1)).astype(np.float32) / np.float32(INTER_TAB_SIZE)
    one = np.float32(1.0)
    w = [(one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx]

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W) & ~bad
        v = A[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[..., None], v, np.float32(0.0))

    out = tap(y0, x0) * w[0][..., None]
    out = out + tap(y0, x0 + 1) * w[1][..., None]
    out = out + tap(y0 + 1, x0) * w[2][..., None]
    out = out + tap(y0 + 1, x0 + 1) * w[3][..., None]
    out = out.astype(np.float32)
    return out[..., 0] if squeeze else out


def sample_maps(flow):
    """The maps the reference builds (coco-generation.py:89-92): float64 sums cast to the image dtype (fp32)."""
    h, w = flow.shape[:2]
    x = (flow[..., 0] + np.arange(w)).astype(np.float32)
    y = (flow[..., 1] + np.arange(h)[:, np.newaxis]).astype(np.float32)
    return x, y


def warp_flow(A, flow):
    """coco-generation.py:86-94 (also ``warp_image`` :66-84 on an (H,W,C) image)."""
    x, y = sample_maps(np.asarray(flow, np.float32))
    return remap_linear(A, x, y)


def _sqnorm2(a, b):
    """``np.linalg.norm([a, b], axis)**2.0`` in fp32: sqrt of the rounded sum of rounded squares, squared again."""
    s = np.sqrt(a * a + b * b, dtype=np.float32)
    return s * s


def gradient_np(f):
    """``np.gradient`` of an (H,W) fp32 plane -> (d/dy, d/dx): halved central differences, one-sided at the borders."""
    f = np.asarray(f, np.float32)
    gy, gx = np.empty_like(f), np.empty_like(f)
    gy[1:-1] = (f[2:] - f[:-2]) / np.float32(2.0)
    gy[0], gy[-1] = f[1] - f[0], f[-1] - f[-2]
    gx[:, 1:-1] = (f[:, 2:] - f[:, :-2]) / np.float32(2.0)
    gx[:, 0], gx[:, -1] = f[:, 1] - f[:, 0], f[:, -1] - f[:, -2]
    return gy, gx


def fb_check(w_warp, w_back, motion_boundaries=True, margins=False):
    """coco-generation.py:96-113 (``motion_boundaries=False``: the COCO copy comments that test out, :111) and
    hollywood2-generation.py:63-81 / sintel-generation.py:89-107 (both tests).  (H,W,2) fp32 each -> (H,W) in {0,1}."""
    w_warp, w_back = np.asarray(w_warp, np.float32), np.asarray(w_back, np.float32)
    s = w_warp + w_back
    norm_wb = _sqnorm2(s[..., 0], s[..., 1])
    norm_w = _sqnorm2(w_warp[..., 0], w_warp[..., 1])
    norm_b = _sqnorm2(w_back[..., 0], w_back[..., 1])
    rhs_occ = np.float32(0.01) * (norm_w + norm_b) + np.float32(0.5)
    occ = norm_wb > rhs_occ
    gy, gx = gradient_np(w_back[..., 0])
    norm_u = _sqnorm2(gy, gx)
    gy, gx = gradient_np(w_back[..., 1])
    norm_v = _sqnorm2(gy, gx)
    rhs_mob = np.float32(0.01) * norm_b + np.float32(0.002)
    mob = (norm_u + norm_v) > rhs_mob
    weights = np.ones(w_warp.shape[:2], np.float32)
    weights[occ] = 0
    if motion_boundaries:
        weights[mob] = 0
    if margins:
        return weights, norm_wb - rhs_occ, (norm_u + norm_v) - rhs_mob
    return weights
