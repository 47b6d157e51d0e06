"""Timing of the training-loss backward kernel alone (tuning aid).  usage: TCL_B200_LIB=... bwd_probe.py"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
dev = torch.device("cuda:0")
pairs = 64
cfg = tcl.synth.CONFIGS["sintel_full"]
ff, bf = tcl.synth.make_flows(pairs, cfg["H"], cfg["W"], seed=5, max_shift=32.0, max_rot_deg=3.0, device=dev)
prev, cur = tcl.synth.make_frames(pairs, 3, cfg["H"], cfg["W"], seed=5, device=dev)
m = tcl.fbcCheckTorch(ff, bf)
gp, gc = torch.empty_like(prev), torch.empty_like(cur)
scale = torch.full((1,), 1e-6, device=dev)
lib = tcl._cabi.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def fn():
    tcl._cabi.check(lib.tclb200_tcl_backward(bf.data_ptr(), m.data_ptr(), prev.data_ptr(), cur.data_ptr(), scale.data_ptr(),
                                             gp.data_ptr(), gc.data_ptr(), pairs, 3, cfg["H"], cfg["W"], 0, 0, st))
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    fn()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
px = pairs * cfg["H"] * cfg["W"]
print(f"tcl_backward (memset + kernel) {ms*1e3:8.1f} us  {px/ms/1e6:6.1f} Gpix/s  [{os.path.basename(os.environ.get('TCL_B200_LIB','default'))}]")
