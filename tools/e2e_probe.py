"""Host-buffer evaluation rate vs the box's concurrent pinned-H2D ceiling, for a few chunk sizes (run under torchrun).
usage: torchrun --nproc-per-node N tools/e2e_probe.py [pairs_per_rank]"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tcl_b200 as tcl  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
device = torch.device("cuda", local)
torch.cuda.set_device(device)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 400
cfg = tcl.synth.CONFIGS["sintel_full"]
H, W = cfg["H"], cfg["W"]
if world > 1:
    bench.bind_to_gpu_numa_node(local)
frames_h = torch.empty((P + 1, 3, H, W), dtype=torch.float32, pin_memory=True)
ff_h = torch.empty((P, 2, H, W), dtype=torch.float32, pin_memory=True)
bf_h = torch.empty((P, 2, H, W), dtype=torch.float32, pin_memory=True)
for s in range(0, P, 32):
    n = min(32, P - s)
    ff, bf = tcl.synth.make_flows(n, H, W, seed=900 + s + 1000 * rank, device=device)
    fr, _ = tcl.synth.make_frames(n + 1, 3, H, W, seed=900 + s, device=device)
    ff_h[s:s + n].copy_(ff); bf_h[s:s + n].copy_(bf); frames_h[s:s + n + 1].copy_(fr)
torch.cuda.synchronize()
pi, ci = torch.arange(0, P, dtype=torch.int32), torch.arange(1, P + 1, dtype=torch.int32)
seq = torch.zeros(P, dtype=torch.long)
ceil_local, first_local = bench.h2d_ceiling(device, [frames_h, ff_h, bf_h], dist if world > 1 else None)
c = torch.tensor([ceil_local, first_local], device=device, dtype=torch.float64)
if world > 1:
    dist.all_reduce(c)
bytes_step = (P + 1) * 3 * H * W * 4 + 2 * P * 2 * H * W * 4
for chunk in (32, 0):
    for _ in range(2):
        tcl.evaluate_sharded_host(frames_h, ff_h, bf_h, pi, ci, seq, 1, chunk_pairs=chunk)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        r = tcl.evaluate_sharded_host(frames_h, ff_h, bf_h, pi, ci, seq, 1, chunk_pairs=chunk)
        float(r["mean_over_pairs"])
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / 4], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gbs = bytes_step * world / float(t[0]) / 1e9
        print(f"N={world} chunk={chunk:4d}: {P * world / float(t[0]):9.0f} pairs/s  {gbs:7.1f} GB/s H2D  = {gbs / float(c[0]):.3f} of the ceiling {float(c[0]):.1f} GB/s (first 2 GiB alone: {float(c[1]):.1f})", flush=True)

# the same bytes in the pipeline's copy pattern (per chunk: a run of frames, the ff chunk, the bf chunk), plain torch copies,
# no kernels and no events: what the pattern alone costs against the big-copy ceiling
def pattern(chunk, streams):
    st = [torch.cuda.Stream(device) for _ in range(streams)]
    d_fr = torch.empty((chunk + 1, 3, H, W), device=device)
    d_f = [torch.empty((chunk, 2, H, W), device=device) for _ in range(2 * streams)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k, s0 in enumerate(range(0, P, chunk)):
        n = min(chunk, P - s0)
        with torch.cuda.stream(st[k % streams]):
            d_fr[:n].copy_(frames_h[s0 + 1:s0 + 1 + n], non_blocking=True)
            d_f[2 * (k % streams)][:n].copy_(ff_h[s0:s0 + n], non_blocking=True)
            d_f[2 * (k % streams) + 1][:n].copy_(bf_h[s0:s0 + n], non_blocking=True)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gbs = bytes_step * world / float(t[0]) / 1e9
        print(f"N={world} copy pattern only, chunk={chunk} streams={streams}: {gbs:7.1f} GB/s = {gbs / float(c[0]):.3f} of the ceiling", flush=True)


for chunk, streams in ((32, 1), (32, 2), (75, 1), (75, 2)):
    pattern(chunk, streams)
    pattern(chunk, streams)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
