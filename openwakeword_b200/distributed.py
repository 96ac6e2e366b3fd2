"""Multi-GPU sharding of independent audio streams / clips.

A stream's scores depend only on its own PCM and its own rings, and the weights are a few MB, so
the path shards with NO data-path collective: rank r owns the contiguous block
``shard_range(n, r, world)`` of streams (or clips), weights are replicated, and the only exchange
is one gather of ``float32[n_local, n_labels]`` per step (``gather_scores``: NCCL all-gather over
NVLink/NVSwitch on GPU ranks, gloo in the CPU tests).  One process per GPU (torchrun).
"""
import os


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank ``rank``; the first n_items % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """torchrun-style init from RANK/WORLD_SIZE/MASTER_*; returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_scores(local_scores, n_total, group=None):
    """All-gather per-rank score blocks [n_local, L] -> [n_total, L] in stream order on every rank.
    Blocks may differ by one row (uneven shard); they are padded to the largest for the collective."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_scores
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    L = local_scores.shape[1]
    pad = local_scores
    if local_scores.shape[0] < n_max:
        pad = torch.zeros((n_max, L), dtype=local_scores.dtype, device=local_scores.device)
        pad[:local_scores.shape[0]] = local_scores
    out = torch.empty((world * n_max, L), dtype=local_scores.dtype, device=local_scores.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(hi - lo == n_max for lo, hi in sizes):
        return out
    return torch.cat([out[r * n_max:r * n_max + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


class ShardedStreams:
    """n_total independent streams split over the ranks of the default process group.

    ``engine_factory(n_local, lo, hi)`` builds this rank's step engine (on GPU ranks a
    ``StreamEngine`` bound to LOCAL_RANK); ``step(local_pcm)`` runs it and gathers the scores."""

    def __init__(self, n_total, engine_factory, rank=None, world=None):
        r, w, _ = env_rank_world()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.n_total = n_total
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        self.engine = engine_factory(self.hi - self.lo, self.lo, self.hi)

    def step(self, local_pcm, n_chunks=1):
        return gather_scores(self.engine.step(local_pcm, n_chunks), self.n_total)
