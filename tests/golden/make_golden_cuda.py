"""Run on a B200 (gpurun): settle which fp32 operation order torch-CUDA uses for this path, write the
CUDA-flavour golden vectors, and cross-check the product kernels.

1. oracle/torch_port.py (bit-identical to the reference on CPU, tests/test_oracle_pinning.py) is run
   on CUDA -> what the reference computes on this GPU.
2. The C oracle is evaluated for all 64 variant masks; the masks that reproduce torch-CUDA bit for bit
   (warp values, mask, both test margins) are reported.  ORACLE_ATEN_CUDA must be among them.
3. Small cases are saved as tests/golden/ref_cuda_*.npz candidates (into gpurun_out/golden_cuda/).
4. The product kernels are compared with torch-CUDA on the same inputs (bitwise counts).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import torch_port as tp  # noqa: E402
import tcl_b200  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(os.path.join(OUT, "golden_cuda"), exist_ok=True)
dev = torch.device("cuda:0")
report = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "cases": []}


def run_case(name, B, H, W, kind, seed, save, **kw):
    ff, bf = tcl_b200.synth.make_flows(B, H, W, seed=seed, **kw)
    prev, cur = tcl_b200.synth.make_frames(B, 3, H, W, seed=seed, kind=kind)
    dff, dbf, dprev, dcur = (t.to(dev) for t in (ff, bf, prev, cur))
    with torch.no_grad():
        t_warp = tp.backward_warp(dprev, dbf)
        t_mask, t_mo, t_mm = tp.fb_consistency(dff, dbf, return_margins=True)
        t_fsw = tp.validity_warp(dprev, dbf)
        t_grad = tp.central_diff(dbf[:, 0])
        t_rmse = tp.tcl_rmse(t_mask, dcur, t_warp)
        t_rmse_ps = tp.tcl_rmse_per_sample(t_mask, dcur, t_warp)
        t_l1 = tp.tcl_l1(t_mask, dcur, t_warp)
    tw, tm, tmo, tmm = t_warp.cpu().numpy(), t_mask.cpu().numpy(), t_mo.cpu().numpy(), t_mm.cpu().numpy()
    matches = []
    stats = {}
    for v in range(64):
        w = oracle.warp(prev.numpy(), bf.numpy(), v)
        m, mo, mm = oracle.fbcheck(ff.numpy(), bf.numpy(), variant=v, margins=True)
        s = dict(warp_ne=int((w != tw).sum()), warp_maxabs=float(np.abs(w - tw).max()), mask_ne=int((m != tm).sum()),
                 mocc_ne=int((mo != tmo).sum()), mmob_ne=int((mm != tmm).sum()))
        stats[v] = s
        if s["warp_ne"] == 0 and s["mask_ne"] == 0 and s["mocc_ne"] == 0 and s["mmob_ne"] == 0:
            matches.append(v)
    # product kernels vs torch-CUDA
    k_warp = tcl_b200.warp(dprev, dbf)
    k_mask, k_near = tcl_b200.fbcheck_with_near_count(dff, dbf)
    k_fsw = tcl_b200.fs_warp(dprev, dbf)
    k_grad = tcl_b200.gradient(dbf[:, 0].contiguous())
    res = tcl_b200.fused_forward(dbf, dprev, dcur, ff=dff)
    near_ref = int(((t_mo.abs() < 1e-6) | (t_mm.abs() < 1e-6)).sum())
    case = dict(name=name, shape=[B, H, W], kind=kind, keep=float(tm.mean()), matching_variants=matches,
                aten_cuda_variant=oracle.ATEN_CUDA, aten_cuda_stats=stats[oracle.ATEN_CUDA],
                best=sorted(((s["warp_ne"] + s["mask_ne"] + s["mocc_ne"] + s["mmob_ne"], v) for v, s in stats.items()))[:5],
                kernel=dict(warp_ne=int((k_warp != t_warp).sum()), warp_maxabs=float((k_warp - t_warp).abs().max()),
                            mask_ne=int((k_mask != t_mask).sum()), fs_warp_ne=int((k_fsw != t_fsw).sum()),
                            grad_ne=int((k_grad != t_grad).sum()), near_kernel=int(k_near), near_torch=near_ref,
                            rmse_kernel=float(res.total_val), rmse_torch=float(t_rmse),
                            rmse_rel=float(abs(float(res.total_val) - float(t_rmse)) / max(float(t_rmse), 1e-30)),
                            rmse_ps_maxrel=float(((res.pair_vals - t_rmse_ps).abs() / t_rmse_ps.clamp(min=1e-30)).max())))
    report["cases"].append(case)
    print(json.dumps(case))
    if save:
        np.savez_compressed(os.path.join(OUT, "golden_cuda", f"ref_cuda_{name}.npz"), ff=ff.numpy(), bf=bf.numpy(),
                            prev=prev.numpy(), cur=cur.numpy(), warp=tw, mask=tm, margin_occ=tmo, margin_mob=tmm,
                            fs_warp=t_fsw.cpu().numpy(), grad_u=t_grad.cpu().numpy(), rmse=t_rmse.cpu().numpy(),
                            rmse_per_sample=t_rmse_ps.cpu().numpy(), l1=t_l1.cpu().numpy(),
                            torch_version=np.array(torch.__version__), gpu=np.array(torch.cuda.get_device_name(0)))


run_case("smooth", 2, 24, 40, "smooth", 100, True, max_shift=5.0, max_rot_deg=4.0, n_rects=3, rect_shift=3.0)
run_case("white_odd", 2, 19, 23, "white", 101, True, max_shift=6.0, max_rot_deg=2.0, n_rects=1, rect_shift=2.5)
run_case("large_disp", 1, 32, 48, "white", 102, True, max_shift=30.0, max_rot_deg=3.0, n_rects=2, rect_shift=12.0)
run_case("mid_white", 2, 128, 160, "white", 103, False, max_shift=10.0)
run_case("sintel_white", 1, 436, 1024, "white", 104, False, max_shift=32.0)
run_case("hd_white", 1, 1080, 1920, "white", 105, False, max_shift=56.0)
with open(os.path.join(OUT, "probe_semantics.json"), "w") as f:
    json.dump(report, f, indent=1)
ok = all(oracle.ATEN_CUDA in c["matching_variants"] for c in report["cases"])
print("ORACLE_ATEN_CUDA reproduces torch-CUDA bitwise on every case:", ok)
