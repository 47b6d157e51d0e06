"""One small call of every kernel of the library, for `compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck`
(the path's equivalent of the race detection the reference does not have):

    compute-sanitizer --tool memcheck python tools/sanitize_target.py

(On this round's GPU pool compute-sanitizer is administratively closed -- "runs under it have left GPUs needing a reset" --
so the script was only run bare, as an every-kernel smoke; the bounds are covered by the parity tests on ragged shapes.)

Shapes are chosen so that edge tiles, mixed tiles, the dynamic tile schedule (> 16 tiles per CTA is not reachable at
sanitizer speed, so the static one), the generic kernel, clip mode, the host pipeline and the backward all run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl

d = torch.device("cuda:0")
B, H, W = 2, 100, 200
ff, bf = tcl.synth.make_flows(B, H, W, seed=3, max_shift=12.0, max_rot_deg=3.0, n_rects=4, rect_shift=10.0, device=d)
prev, cur = tcl.synth.make_frames(B, 3, H, W, seed=3, kind="white", device=d)
r = tcl.fused_forward(bf, prev, cur, ff=ff)                                         # hot kernel (edge + mixed tiles)
r2 = tcl.fused_forward(bf, prev, cur, ff=ff, want_warp=True, want_mask=True, want_blend=True)   # outputs variant
r3 = tcl.fused_forward(bf, prev, cur, ff=ff, want_warp=True, want_mask=True, want_near=True)    # feature-complete path
m = tcl.fbcCheckTorch(ff, bf)
tcl.fbcCheckTorch_mob(ff, bf)
tcl.gradient(bf[:, 0])
tcl.warp(prev, bf); tcl.fs_warp(prev, bf)
tcl.temporal_loss(m, cur, prev, bf, loss="l1")
tcl.warp_blend(m, prev, bf, cur)
tcl.temporal_error(ff, bf, prev.bfloat16(), cur.bfloat16())
p = prev.clone().requires_grad_(True); c = cur.clone().requires_grad_(True); f = bf.clone().requires_grad_(True)
tcl.temporal_loss(m, c, p, bf).backward()
tcl.warp(p, f).sum().backward()
tcl._cabi.lib().tclb200_debug_force_generic(1)
tcl.fused_forward(bf, prev, cur, ff=ff, want_warp=True, want_mask=True)           # generic kernel
tcl.fused_forward(bf[:, :, :, :199].contiguous(), prev[:, :, :, :199].contiguous(), cur[:, :, :, :199].contiguous(), ff=ff[:, :, :, :199].contiguous())
tcl._cabi.lib().tclb200_debug_force_generic(0)
frames = torch.cat([prev[:1], cur], 0)
tcl.temporal_error_clip(frames, ff, bf)
tcl.temporal_error_host(frames.cpu().pin_memory(), ff.cpu().pin_memory(), bf.cpu().pin_memory(), chunk_pairs=1)
tcl.upsample_flow(torch.randn(1, 2, 12, 20, device=d), torch.randn(1, 576, 12, 20, device=d))
tcl.split_fc2_block(torch.randn(2, 32, 48, 9, device=d))
cc = tcl.cv2compat
hw_ff, hw_bf = ff[0].permute(1, 2, 0).contiguous(), bf[0].permute(1, 2, 0).contiguous()
cc.fb_check_flows(hw_ff, hw_bf); cc.warp_image(prev[0].permute(1, 2, 0).contiguous(), hw_bf)
idx = tcl.window_evaluations(3, 3)                                                   # window mode (flow / frame banks, interleaved tiles)
bank = torch.cat([bf, ff, bf[:1], ff[:1]], 0)[: 2 * idx["field_t"].numel()].contiguous()
tcl.temporal_error_window(torch.cat([prev, cur[:1]], 0), bank, 3, idx)
tcl.fused_forward(bf, prev, cur, ff=ff, rows=(17, 71))                               # band mode
tcl.reconet_output_temporal_loss(m, cur, prev, cur * 0.5, prev * 0.5, bf)            # learning-based chains
tcl.ruder_network_input(cur, m, prev, bf)
torch.cuda.synchronize()
print("sanitize_target ok", float(r.total_val), float(r2.total_val), float(r3.total_val))
