"""Probe (not a pytest file): concurrent pinned D2H / H2D bandwidth of all ranks of a torchrun launch.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/d2h_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
n = 2 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.fill_(1)
res = {}
for name, (dst, src) in (("d2h", (host, dev)), ("h2d", (dev, host))):
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    res[name] = 5 * n / (time.perf_counter() - t0) / 1e9
print("rank %d of %d: D2H %.1f GB/s, H2D %.1f GB/s (2 GiB pinned, all ranks at once)" % (rank, world, res["d2h"], res["h2d"]), flush=True)
if world > 1:
    dist.destroy_process_group()
