"""A few launches of every secondary kernel on the Sintel shape (the target of the per-kernel ncu captures in profiles/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl  # noqa: E402

d = torch.device("cuda:0")
cfg = tcl.synth.CONFIGS["sintel_full"]
B, H, W = 32, cfg["H"], cfg["W"]
ff, bf = tcl.synth.make_flows(B, H, W, seed=5, max_shift=cfg["max_shift"], max_rot_deg=cfg["max_rot_deg"], device=d)
prev, cur = tcl.synth.make_frames(B, 3, H, W, seed=5, device=d)
mask = tcl.fbcCheckTorch(ff, bf)
cc = tcl.cv2compat
hw_ff, hw_bf = ff.permute(0, 2, 3, 1).contiguous(), bf.permute(0, 2, 3, 1).contiguous()
hw_img = prev.permute(0, 2, 3, 1).contiguous()
for _ in range(3):
    tcl.upsample_flow(torch.randn(4, 2, H // 8, W // 8, device=d), torch.randn(4, 576, H // 8, W // 8, device=d))
    cc.fb_check_flows(hw_ff, hw_bf)
    cc.warp_image(hw_img, hw_bf)
    tcl.split_fc2_block(torch.randn(B, 256, 256, 9, device=d))
    tcl._cabi.lib().tclb200_debug_force_generic(1)
    tcl.fused_forward(bf, prev, cur, ff=ff)
    tcl._cabi.lib().tclb200_debug_force_generic(0)
    p = prev.clone().requires_grad_(True)
    tcl.temporal_loss(mask, cur, p, bf).backward()
    tcl.reconet_output_temporal_loss(mask, cur, prev, cur * 0.5, prev * 0.5, bf)
    tcl.ruder_network_input(cur, mask, prev, bf)
torch.cuda.synchronize()
print("prof_ops ok")
