"""Aggregate host-to-device bandwidth of the node: every rank copies a pinned 256 MB buffer to its GPU in a loop
(nothing else running).  Run under torchrun with N ranks; prints GB/s per rank and in total.  The shard loader's
end-to-end rate at N ranks cannot exceed total / (4 * d_model) activations/s."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 40
t0 = time.perf_counter()
for _ in range(reps):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = torch.tensor([n * reps / dt / 1e9], device="cuda")
if world > 1:
    all_g = [torch.zeros_like(gbs) for _ in range(world)]
    dist.all_gather(all_g, gbs)
    vals = [float(g) for g in all_g]
else:
    vals = [float(gbs)]
if rank == 0:
    print(f"H2D pinned copy, {world} rank(s) concurrently: per rank {[round(v, 1) for v in vals]} GB/s, total {sum(vals):.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
