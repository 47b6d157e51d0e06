"""Small fixed workload for ncu: N launches of the fused kernel on a Sintel-shape shard (inputs > L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
workload = sys.argv[3] if len(sys.argv) > 3 else "sintel_full"
dev = torch.device("cuda:0")
cfg = tcl.synth.CONFIGS[workload]
dt = torch.bfloat16 if cfg["dtype"] == "bf16" else torch.float32
ff, bf = tcl.synth.make_flows(pairs, cfg["H"], cfg["W"], seed=3234, max_shift=cfg["max_shift"], max_rot_deg=cfg["max_rot_deg"], device=dev)
prev, cur = tcl.synth.make_frames(pairs, 3, cfg["H"], cfg["W"], seed=3234, kind="smooth", device=dev, dtype=dt)
torch.cuda.synchronize()
for _ in range(launches):
    r = tcl.fused_forward(bf, prev, cur, ff=ff)
torch.cuda.synchronize()
print("rmse", float(r.total_val), "pairs", pairs)
