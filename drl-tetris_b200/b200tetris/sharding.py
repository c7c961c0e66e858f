"""Multi-GPU plumbing: games are independent, so N GPUs = N handles on contiguous ranges of GLOBAL game ids and no
collective in the step path.  Seeds and synthetic actions are keyed by the global id, so every game's
trajectory is the same whatever the number of ranks.  The only communication is the optional end-of-rollout sum of the
episode counters (torch.distributed all_reduce; NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np

STAT_KEYS = ("env_steps", "episodes", "wins_p0", "wins_p1", "both_dead", "lines", "garbage_queue_overflow", "capacity_overflow")


def shard_range(n_total, rank, world):
    """Contiguous range [lo, hi) of global game ids owned by `rank` (game g -> rank g*world//n_total)."""
    lo = (n_total * rank + world - 1) // world
    hi = (n_total * (rank + 1) + world - 1) // world
    return lo, hi


def global_ids(n_total, rank, world):
    lo, hi = shard_range(n_total, rank, world)
    return np.arange(lo, hi, dtype=np.int64)


def owner_of(game, n_total, world):
    return int(game) * world // n_total


def reduce_stats(stats, device=None):
    """Sum a stats dict (Engine.stats()) over all ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    t = torch.tensor([int(stats[k]) for k in STAT_KEYS], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return dict(zip(STAT_KEYS, (int(x) for x in t.cpu())))
