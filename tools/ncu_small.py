"""ncu target: the training-batch call (16 x 256 x 256, dataset mask)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ff, bf = tcl.synth.make_flows(B, 256, 256, seed=5, max_shift=24.0, max_rot_deg=6.0, device=dev)
prev, cur = tcl.synth.make_frames(B, 3, 256, 256, seed=5, device=dev)
m = tcl.fbcCheckTorch(ff, bf)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(4):
    r = tcl.fused_forward(bf, prev, cur, mask=m, finalize=tcl.ops.FIN_MEAN)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(float(r.total_val))
