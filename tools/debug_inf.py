"""Which rank's weak shard of hd1080_window gives a non-finite aggregate (seen once at N = 8)? Per-pair values of every path."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
import bench
dev = torch.device("cuda:0")
lib = tcl._cabi.lib()
for rank in range(8):
    sh = bench.make_shard(tcl, "hd1080_window", 6, 1234 + 2000 + 100000 * rank, dev, "smooth")
    fin = {k: bool(torch.isfinite(v.float()).all()) for k, v in sh.items()}
    r = tcl.fused_forward(sh["bf"], sh["prev"], sh["cur"], ff=sh["ff"])
    lib.tclb200_debug_force_generic(1)
    g = tcl.fused_forward(sh["bf"], sh["prev"], sh["cur"], ff=sh["ff"])
    lib.tclb200_debug_force_generic(0)
    seq = torch.zeros(6, dtype=torch.long, device=dev)
    ev = tcl.evaluate_sharded(sh["ff"], sh["bf"], sh["prev"], sh["cur"], seq, 1)
    print(rank, fin, "hot", [f"{v:.6f}" for v in r.pair_vals.tolist()], "generic", [f"{v:.6f}" for v in g.pair_vals.tolist()],
          "agg", float(ev["mean_over_sequences"]), float(ev["pooled_rmse"]), "flow max", float(sh["bf"].abs().max()), float(sh["ff"].abs().max()), flush=True)
