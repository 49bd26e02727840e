#!/usr/bin/env python3
"""Raw pinned-host -> device copy rate with every rank of a torchrun job copying at once (the ceiling of the e2e leg)."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 4 << 30
host = torch.empty(size, dtype=torch.uint8).pin_memory()
host.fill_(7)
dev = torch.empty(size, dtype=torch.uint8, device="cuda")
best = 0.0
for _ in range(4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    best = max(best, size * world / float(dt.item()) / 1e9)
if rank == 0:
    print(f'{{"probe": "raw pinned H2D, all ranks at once", "n_gpus": {world}, "gib_per_gpu": {size / (1 << 30):.1f}, "aggregate_gbs": {best:.1f}, "per_gpu_gbs": {best / world:.1f}, "cpus": {os.cpu_count()}}}')
if world > 1:
    dist.destroy_process_group()
