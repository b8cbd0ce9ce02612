"""CPU-only: the multi-GPU batch path's host logic at world_size 2 over gloo:
round-robin frame sharding with no data-path collective, whole-job rate = all units / max-over-ranks time."""
import importlib
import os
import socket
import sys

import pytest

from conftest import PKG, ROOT


def test_round_robin_partition():
    shard = importlib.import_module(PKG + ".shard")
    for n in (0, 1, 7, 8, 1024):
        for world in (1, 2, 4, 8):
            parts = [shard.frames_for_rank(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            sizes = shard.shard_sizes(n, world)
            assert sizes == [len(p) for p in parts] and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.frames_for_rank(4, 2, 2)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    shard = importlib.import_module(PKG + ".shard")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.frames_for_rank(11, rank, world)
    # every rank "processes" its frames; rank 1 is slower: the job rate must use the max time and all units
    secs = 2.0 if rank == 1 else 1.0
    rate, units, t = shard.job_rate(dist, len(mine), secs)
    got = [None] * world
    dist.all_gather_object(got, mine)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, rate, units, t, got))


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rate, units, t, got in res:
        assert units == 11 and t == 2.0 and abs(rate - 5.5) < 1e-12
        assert sorted(got[0] + got[1]) == list(range(11)) and not set(got[0]) & set(got[1])


def test_strip_bands():
    """row-strip mode: bands are whole tile rows, cover the image once, and every inner side gets two halo rows"""
    strips = importlib.import_module(PKG + ".strips")
    for h in (32, 1080, 2160, 30000):
        for world in (1, 2, 4, 8):
            bands = strips.row_bands(h, world)
            assert bands[0][0] == 0 and bands[-1][1] == h
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0
            for a0, a1 in bands:
                assert a0 % 32 == 0 and (a1 % 32 == 0 or a1 == h)
                r0, r1 = strips.halo_rows(a0, a1, h)
                assert (a0 == 0 or a0 - r0 == 2) and (a1 == h or r1 - a1 == 2)
