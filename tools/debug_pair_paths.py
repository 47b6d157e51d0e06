"""Localise the non-finite per-pair value of the bf16 hot path (hd1080_window, weak shard of rank 7, pair 2)."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
import bench
dev = torch.device("cuda:0")
lib = tcl._cabi.lib()
sh = bench.make_shard(tcl, "hd1080_window", 6, 1234 + 2000 + 100000 * 7, dev, "smooth")
ff, bf, prev, cur = (sh[k][2:3].contiguous() for k in ("ff", "bf", "prev", "cur"))
H, W = 1080, 1920
def stats():
    st = (ctypes.c_ulonglong * 2)()
    lib.tclb200_debug_tile_stats(st, 1)
    return (st[0], st[1])
lib.tclb200_debug_tile_stats(None, 1)
r = tcl.fused_forward(bf, prev, cur, ff=ff)
print("single pair hot bf16:", r.pair_sums.tolist(), "global/mixed tiles", stats())
r32 = tcl.fused_forward(bf, prev.float(), cur.float(), ff=ff)
print("single pair hot fp32 frames:", r32.pair_sums.tolist(), stats())
r4 = tcl.fused_forward(bf, prev, cur, ff=ff, want_warp=True, want_mask=True)
print("outputs path bf16:", r4.pair_sums.tolist(), "warp finite", bool(torch.isfinite(r4.warp.float()).all()), stats())
lib.tclb200_debug_force_generic(1)
g = tcl.fused_forward(bf, prev, cur, ff=ff, want_warp=True, want_mask=True)
lib.tclb200_debug_force_generic(0)
print("generic:", g.pair_sums.tolist())
bad = []
for r0 in range(0, H, 32):
    r1 = min(H, r0 + 32)
    b = tcl.fused_forward(bf, prev, cur, ff=ff, rows=(r0, r1))
    s = stats()
    lib.tclb200_debug_force_generic(1)
    gb = tcl.fused_forward(bf, prev, cur, ff=ff, rows=(r0, r1))
    lib.tclb200_debug_force_generic(0)
    stats()
    v, gv = float(b.pair_sums[0]), float(gb.pair_sums[0])
    if not (abs(v - gv) <= 1e-4 * abs(gv)):
        bad.append(r0)
        print("band", r0, r1, "hot", v, "generic", gv, "global/mixed", s)
# per-tile extents of the bad bands (what the scanner sees)
xs = torch.arange(W, device=dev, dtype=torch.float32)[None, :]
ys = torch.arange(H, device=dev, dtype=torch.float32)[:, None]
ax, ay = xs + bf[0, 0], ys + bf[0, 1]
i2x, i2y = 2.0 * (1.0 / (W - 1)), 2.0 * (1.0 / (H - 1))
cx = (((ax * i2x - 1) + 1) * W - 1) * 0.5
cy = (((ay * i2y - 1) + 1) * H - 1) * 0.5
for r0 in bad[:4]:
    for c0 in range(0, W, 64):
        tx, ty = cx[r0:r0 + 32, c0:c0 + 64], cy[r0:r0 + 32, c0:c0 + 64]
        bx0, bx1, by0, by1 = int(tx.min().floor()), int(tx.max().floor()), int(ty.min().floor()), int(ty.max().floor())
        ox = bx0 & ~7
        fitx, fity = bx1 + 1 - ox < 80, by1 + 1 - by0 < 42
        if not (fitx and fity) or bx0 < 0 or by0 < 0 or bx1 + 1 >= W or by1 + 1 >= H:
            print(f"  tile row {r0} col {c0}: x [{bx0},{bx1}] y [{by0},{by1}] ox {ox} fitx {fitx} fity {fity}")
