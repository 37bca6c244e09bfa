"""CPU: the N>1 host logic (image sharding + max-over-ranks reduction) with a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ocljpegdecoder_b200.sharding import shard_by_bytes, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 256, 8192):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_bytes_is_contiguous_and_complete():
    sizes = [5, 1, 1, 9, 2, 2, 2, 7, 3]
    for world in (1, 2, 3, 4):
        spans = shard_by_bytes(sizes, world)
        assert spans[0][0] == 0 and spans[-1][1] == len(sizes)
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        assert all(hi > lo for lo, hi in spans)


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_items, rank, world)
    owned = torch.zeros(n_items, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)                       # every image is decoded by exactly one rank
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # the timing rule: max over ranks
    pix = torch.tensor([float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(pix)                         # whole-job units
    if rank == 0:
        q.put((owned.tolist(), float(t.item()), float(pix.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_items = 257
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    owned, tmax, units = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert owned == [1] * n_items and tmax == 2.0 and units == n_items
