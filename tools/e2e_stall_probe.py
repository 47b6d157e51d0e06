"""Which call of a short host-entry step stalls?  N = 1, hd1080_window: 400 steps with a timer around every piece."""
import os, sys, time, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
from tcl_b200 import ops, sharding, _cabi
import bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
wl = "hd1080_window"
cfg = tcl.synth.CONFIGS[wl]; n = cfg["pairs"]
sh = bench.make_shard(tcl, wl, n, 3234, dev, "smooth")
F = n + 1
frames_h = torch.empty((F,) + tuple(sh["cur"].shape[1:]), dtype=sh["cur"].dtype, pin_memory=True)
ff_h = torch.empty(sh["ff"].shape, dtype=torch.float32, pin_memory=True); bf_h = torch.empty(sh["ff"].shape, dtype=torch.float32, pin_memory=True)
ff_h.copy_(sh["ff"]); bf_h.copy_(sh["bf"]); frames_h[0].copy_(sh["prev"][0]); frames_h[1:].copy_(sh["cur"])
pi, ci = torch.arange(0, n, dtype=torch.int32), torch.arange(1, n + 1, dtype=torch.int32)
seq = torch.zeros(n, dtype=torch.long)
T = {}
def tick(name, t0):
    t1 = time.perf_counter(); T.setdefault(name, []).append((t1 - t0) * 1e3); return t1
orig_mem, orig_empty = torch.cuda.mem_get_info, torch.empty
def mem(*a, **k):
    t0 = time.perf_counter(); r = orig_mem(*a, **k); tick("mem_get_info", t0); return r
torch.cuda.mem_get_info = mem
lib = _cabi.lib()
orig_host = lib.tclb200_tcl_forward_host
class L:   # proxy that times the C call
    def __getattr__(self, k): return getattr(lib, k)
    def tclb200_tcl_forward_host(self, *a):
        t0 = time.perf_counter(); r = orig_host(*a); tick("C call (enqueue)", t0); return r
_cabi_lib = _cabi.lib
_cabi.lib = lambda: L()
orig_sync = torch.cuda.Stream.synchronize
def sync(self):
    t0 = time.perf_counter(); r = orig_sync(self); tick("stream.synchronize", t0); return r
torch.cuda.Stream.synchronize = sync
for i in range(400):
    t0 = time.perf_counter()
    vals, sums = ops.temporal_error_host(frames_h, ff_h, bf_h, pi, ci, device=dev, return_sums=True)
    t1 = tick("temporal_error_host total", t0)
    packed = sharding.pack_local(vals.to(dev, non_blocking=True), sums.to(dev, non_blocking=True).sum(), torch.as_tensor(seq).to(dev), 1, 3 * 1080 * 1920)
    t2 = tick("pack_local (+3 small copies)", t1)
    r = sharding.unpack(sharding.allreduce_sums(packed), 1)
    t3 = tick("unpack", t2)
    float(r["mean_over_pairs"])
    tick("float() sync", t3)
    tick("step", t0)
for k, v in T.items():
    n_calls = len(v)
    v = sorted(v[20:]) or sorted(v)
    print(f"{k:32s} median {v[len(v)//2]:8.3f}  p90 {v[int(len(v)*0.9)]:8.3f}  p99 {v[int(len(v)*0.99)]:8.3f}  max {v[-1]:8.3f} ms  ({n_calls} calls)", flush=True)
