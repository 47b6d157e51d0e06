"""Where does a short host-entry step go?  hd1080_window (6 bf16 pairs, 286 MB per rank and step) under torchrun:
plain copies of the same bytes vs temporal_error_host vs evaluate_sharded_host, with / without the NUMA binding of bench.py.
usage: torchrun --nproc-per-node N tools/e2e_small_probe.py [workload]"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcl_b200 as tcl
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "hd1080_window"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = tcl.synth.CONFIGS[wl]
n = cfg["pairs"]
sh = bench.make_shard(tcl, wl, n, 1234 + 2000 + 100000 * rank, dev, "smooth")

def maxr(v):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t[0])

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def run(tag):
    F = n + 1
    frames_h = torch.empty((F,) + tuple(sh["cur"].shape[1:]), dtype=sh["cur"].dtype, pin_memory=True)
    ff_h = torch.empty(sh["ff"].shape, dtype=torch.float32, pin_memory=True); bf_h = torch.empty_like(ff_h).pin_memory()
    ff_h.copy_(sh["ff"]); bf_h.copy_(sh["bf"]); frames_h[0].copy_(sh["prev"][0]); frames_h[1:].copy_(sh["cur"])
    pi, ci = torch.arange(0, n, dtype=torch.int32), torch.arange(1, n + 1, dtype=torch.int32)
    seq = torch.zeros(n, dtype=torch.long)
    d_fr, d_ff, d_bf = torch.empty_like(frames_h, device=dev), torch.empty_like(ff_h, device=dev), torch.empty_like(bf_h, device=dev)
    nbytes = frames_h.numel() * frames_h.element_size() + 2 * ff_h.numel() * 4
    def t_copy():
        d_fr.copy_(frames_h, non_blocking=True); d_ff.copy_(ff_h, non_blocking=True); d_bf.copy_(bf_h, non_blocking=True); torch.cuda.synchronize()
    def t_host():
        tcl.temporal_error_host(frames_h, ff_h, bf_h, pi, ci, device=dev)
    def t_host_ring():
        tcl.temporal_error_host(frames_h, ff_h, bf_h, pi, ci, device=dev, max_device_frames=F)
    def t_host_c2():
        tcl.temporal_error_host(frames_h, ff_h, bf_h, pi, ci, device=dev, chunk_pairs=2)
    def t_eval():
        float(tcl.evaluate_sharded_host(frames_h, ff_h, bf_h, pi, ci, seq, 1)["mean_over_pairs"])
    for name, fn in (("plain copies", t_copy), ("temporal_error_host", t_host), ("... max_device_frames=F (no mem_get_info)", t_host_ring),
                     ("... chunk_pairs=2", t_host_c2), ("evaluate_sharded_host", t_eval)):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        ms = maxr((time.perf_counter() - t0) / 10 * 1e3)
        if rank == 0:
            print(f"[{tag}] N={world} {wl}: {name:45s} {ms:8.2f} ms/step  {nbytes * world / ms / 1e6:7.1f} GB/s aggregate", flush=True)

run("unbound")
saved = os.sched_getaffinity(0)
bench.bind_to_gpu_numa_node(dev.index)
if rank == 0:
    print("affinity", len(saved), "->", len(os.sched_getaffinity(0)), flush=True)
run("bound to the GPU's NUMA node")
