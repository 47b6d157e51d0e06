"""Isolate kernel paths in separate processes (a faulting kernel kills the CUDA context)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    "generic_warp": "tcl._cabi.lib().tclb200_debug_force_generic(1); y=tcl.warp(prev,bf)",
    "generic_fused": "tcl._cabi.lib().tclb200_debug_force_generic(1); y=tcl.fused_forward(bf,prev,cur,ff=ff).total_val",
    "tma_mob_only": "y=tcl.fbcCheckTorch_mob(None,bf)",
    "tma_fbcheck": "y=tcl.fbcCheckTorch(ff,bf)",
    "tma_warp": "y=tcl.warp(prev,bf)",
    "tma_fused": "y=tcl.fused_forward(bf,prev,cur,ff=ff).total_val",
    "gradient": "y=tcl.gradient(bf[:,0].contiguous())",
}
TEMPLATE = """
import sys, torch
sys.path.insert(0, {root!r})
import tcl_b200 as tcl
d = torch.device('cuda:0')
H, W = {H}, {W}
ff, bf = tcl.synth.make_flows(1, H, W, seed=1, max_shift=3.0, device=d)
prev, cur = tcl.synth.make_frames(1, 3, H, W, seed=1, kind='white', device=d)
torch.cuda.synchronize()
{stmt}
torch.cuda.synchronize()
print('OK', float(y.float().sum()))
"""

if __name__ == "__main__":
    shapes = [(128, 256), (32, 48)]
    for H, W in shapes:
        for name, stmt in CASES.items():
            code = TEMPLATE.format(root=ROOT, H=H, W=W, stmt=stmt)
            env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
            tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or [""])[-1][:200]
            print(f"{H}x{W} {name}: rc={r.returncode} {tail}", flush=True)
