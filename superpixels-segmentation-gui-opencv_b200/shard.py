"""Batch sharding across GPUs (BASELINE.json configs[4]).

Frames are independent, so a batch shards with NO data-path collective: frame i belongs to rank i mod world
(round-robin, so every rank's share differs by at most one frame and a stream of frames stays balanced).
torch.distributed is used only for the barrier and the max-over-ranks timing / result counters.
"""


def frames_for_rank(n_frames, rank, world):
    """indices of the frames rank `rank` of `world` processes"""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard request: n_frames=%r rank=%r world=%r" % (n_frames, rank, world))
    return list(range(rank, n_frames, world))


def shard_sizes(n_frames, world):
    return [len(range(r, n_frames, world)) for r in range(world)]


def job_rate(dist, units_local, seconds_local, device=None):
    """whole-job throughput = units processed by all ranks / max over ranks of the elapsed time.
    `dist` is torch.distributed (initialised) or None for a single process."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return units_local / seconds_local, units_local, seconds_local
    t = torch.tensor([float(seconds_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / float(t.item()), float(u.item()), float(t.item())
