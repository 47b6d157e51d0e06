"""bf16-frame configuration of the hot kernel: burst and sustained throughput on the Sintel and 1080p shapes (tuning aid)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
dev = torch.device("cuda:0")

def run(name, H, W, pairs, shift, rot):
    chunks = []
    for s in range(0, pairs, 16):
        n = min(16, pairs - s)
        ff, bf = tcl.synth.make_flows(n, H, W, seed=77 + s, max_shift=shift, max_rot_deg=rot, device=dev)
        prev, cur = tcl.synth.make_frames(n, 3, H, W, seed=77 + s, device=dev, dtype=torch.bfloat16)
        chunks.append((ff, bf, prev, cur))
    ff, bf, prev, cur = (torch.cat([c[i] for c in chunks]) for i in range(4))
    del chunks
    fn = lambda: tcl.fused_forward(bf, prev, cur, ff=ff)
    import ctypes
    lib = tcl._cabi.lib()
    lib.tclb200_debug_tile_stats(None, 1)
    r = fn()
    st = (ctypes.c_ulonglong * 2)()
    lib.tclb200_debug_tile_stats(st, 1)
    def timed(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    time.sleep(1.0)
    burst = min(timed(3) for _ in range(3))
    t0 = time.time()
    while time.time() - t0 < 1.5:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    sus = timed(40)
    px = pairs * H * W
    print(f"{name:8s} bf16 pairs={pairs:3d} burst {px/burst/1e6:6.1f} Gpix/s ({px*28/burst/1e6/6548.2:.2f})  sustained {px/sus/1e6:6.1f} Gpix/s ({px*28/sus/1e6/6548.2:.2f})"
          f"  mixed tiles {st[1]}/{pairs * ((H + 31) // 32) * ((W + 63) // 64)}  total {float(r.total_val):.9f}  [{os.path.basename(os.environ.get('TCL_B200_LIB', 'product'))}]", flush=True)

run("sintel", 436, 1024, 256, 32.0, 3.0)
run("hd1080", 1080, 1920, 48, 64.0, 3.0)
