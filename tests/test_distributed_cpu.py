"""not-gpu: the N>1 path on CPU with the gloo backend (world_size 2): contiguous stream shards,
no data-path collective, one score all-gather per step (openwakeword_b200/distributed.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openwakeword_b200 import distributed as owd


def test_shard_range_covers_everything_contiguously():
    for n in (0, 1, 7, 8, 1000, 65536, 1_000_000):
        for world in (1, 2, 3, 8):
            spans = [owd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        owd.shard_range(10, 2, 2)


class _FakeEngine:
    """Scores are a deterministic function of the GLOBAL stream id and the step's PCM, so the test can
    tell whether every rank processed its own shard and whether the gather preserved stream order."""

    def __init__(self, n_local, lo, hi):
        self.lo, self.hi, self.n_cols = lo, hi, 3
        self.state = torch.zeros(n_local)

    def step(self, local_pcm, n_chunks=1):
        self.state += local_pcm.float().mean(dim=1)
        ids = torch.arange(self.lo, self.hi, dtype=torch.float32)
        return torch.stack([ids, self.state, ids * 0 + n_chunks], dim=1)


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = owd.init_process_group("gloo")
    assert (r, w) == (rank, world)
    sh = owd.ShardedStreams(n_total, _FakeEngine)
    g = torch.Generator().manual_seed(5)
    out = None
    for step in range(3):
        full = torch.randint(-1000, 1000, (n_total, 1280), generator=g, dtype=torch.int16)   # same on every rank
        out = sh.step(full[sh.lo:sh.hi], 1)
    q.put((rank, out.numpy(), (sh.lo, sh.hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 11])
def test_two_rank_gloo_gather_preserves_stream_order(n_total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, a, span0), (_, b, span1) = res
    assert span0[1] == span1[0] and span0[0] == 0 and span1[1] == n_total
    np.testing.assert_array_equal(a, b)                                   # every rank holds the full result
    assert a.shape == (n_total, 3)
    np.testing.assert_array_equal(a[:, 0], np.arange(n_total))            # stream order preserved across uneven shards
    g = torch.Generator().manual_seed(5)
    acc = torch.zeros(n_total)
    for step in range(3):
        acc += torch.randint(-1000, 1000, (n_total, 1280), generator=g, dtype=torch.int16).float().mean(dim=1)
    np.testing.assert_allclose(a[:, 1], acc.numpy(), rtol=1e-6)


def test_gather_is_identity_without_a_process_group():
    x = torch.arange(12.0).reshape(4, 3)
    assert owd.gather_scores(x, 4) is x
