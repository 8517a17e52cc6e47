"""Pinned host -> device bandwidth per rank when all ranks copy at once (torchrun --nproc-per-node N tools/h2d_probe.py)."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
src = torch.randint(0, 256, (8, 3, 1024, 1024), dtype=torch.uint8).pin_memory()
dst = torch.empty_like(src, device="cuda")
for sz_name, n in (("25 MB x 40 back to back", 40),):
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("rank %d: %s: %.3f ms per copy = %.1f GB/s" % (rank, sz_name, ms, src.numel() / ms / 1e6), flush=True)
