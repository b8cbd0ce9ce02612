"""CPU-only: host logic of the row-strip partition (spx_slic_strip_plan -- no device needed):
the cluster index ranges every rank maintains, the ranges adjacent ranks share, and the refusal of bands too thin for
the neighbour exchange.  The same arithmetic runs inside spx_slic_strip_partition on the GPU box."""
import ctypes as C
import importlib

import pytest

from conftest import PKG


def _plan(spx, w, h, S, bands):
    L = spx.lib()
    L.spx_slic_strip_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    flat = [v for b in bands for v in b]
    own = (C.c_int * len(flat))(*flat)
    out = (C.c_int * (4 * len(bands)))()
    rc = L.spx_slic_strip_plan(w, h, S, own, len(bands), out)
    return rc, [tuple(out[4 * q:4 * q + 4]) for q in range(len(bands))]


@pytest.mark.parametrize("w,h,S,world", [(16384, 16384, 20, 8), (16384, 16384, 20, 4), (8192, 8192, 20, 4), (4096, 4096, 20, 2),
                                         (3840, 2160, 14, 3), (5000, 7001, 33, 5)])
def test_plan_properties(w, h, S, world):
    spx = importlib.import_module(PKG)
    strips = importlib.import_module(PKG + ".strips")
    bands = strips.row_bands(h, world)
    rc, plan = _plan(spx, w, h, S, bands)
    assert rc == 0, spx.lib().spx_last_error()
    xstrips, ystrips = int(0.5 + w / S), int(0.5 + h / S)
    K = xstrips * ystrips
    yerr = int((h - S * ystrips) / ystrips)
    seed_y = lambda gy: min(gy * S + S // 2 + gy * yerr, h - 1)
    # core ranges: a partition of all clusters, in rank order, whole seed rows
    assert plan[0][2] == 0 and plan[-1][3] == K
    for q in range(world):
        lo, hi, clo, chi = plan[q]
        assert lo % xstrips == 0 and hi % xstrips == 0 and lo <= clo <= chi <= hi
        if q:
            assert plan[q - 1][3] == clo
        # every cluster whose window or pixels can reach the band (seed row within 5S rows) is maintained
        for gy in range(ystrips):
            near = bands[q][0] - 5 * S <= seed_y(gy) < bands[q][1] + 5 * S
            assert near == (lo <= gy * xstrips < hi), (q, gy)
    # only adjacent ranks share clusters
    for q in range(world - 2):
        assert plan[q][1] <= plan[q + 2][0]


def test_plan_refuses_thin_bands():
    spx = importlib.import_module(PKG)
    strips = importlib.import_module(PKG + ".strips")
    rc, _ = _plan(spx, 2048, 1536, 20, strips.row_bands(1536, 16))      # 96 rows per rank < 10S + two seed pitches
    assert rc != 0
    rc, _ = _plan(spx, 2048, 1536, 20, [(0, 512), (512, 512), (512, 1536)])   # a rank without rows
    assert rc != 0


def _gloo_worker(rank, world, port, q):
    import os
    import sys
    import torch.distributed as dist
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    spx = importlib.import_module(PKG)
    strips = importlib.import_module(PKG + ".strips")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, S = 8192, 8192, 20
    bands = strips.row_bands(H, world)
    rc, plan = _plan(spx, W, H, S, bands)
    got = [None] * world
    dist.all_gather_object(got, (rc, plan))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, got))


def test_two_ranks_agree_on_the_partition_gloo():
    """world_size 2 over gloo (no GPU): every rank derives the same partition from the image geometry alone -- what
    StripSLIC relies on when each rank calls spx_slic_strip_partition on its own -- and the ranges the two ranks share
    are the same seen from either side"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert out[0] == out[1]
    (rc0, plan0), (rc1, plan1) = out[0]
    assert rc0 == 0 and rc1 == 0 and plan0 == plan1
    (lo0, hi0, _, _), (lo1, hi1, _, _) = plan0
    assert max(lo0, lo1) < min(hi0, hi1)                 # the two ranks share the clusters around the cut ...
    assert lo0 == 0 and hi1 == plan0[1][3]                # ... and together cover every cluster
