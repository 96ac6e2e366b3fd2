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


# ---- PeerGather protocol on host memory: ranks are threads, "device addresses" are addresses of NumPy buffers ----------
class _FakePeerCtx:
    """peer_* calls of _native.Context on host memory, executed synchronously (the real ones are stream-ordered
    kernels; the ordering constraints they must satisfy are the same)."""
    registry = {}

    def __init__(self):
        self.keep = []

    def peer_alloc(self, n_bytes):
        import ctypes
        buf = np.zeros(n_bytes, np.uint8)
        self.keep.append(buf)
        handle = (b"%060d" % len(_FakePeerCtx.registry)).ljust(64, b"\0")
        _FakePeerCtx.registry[handle] = buf
        return buf.ctypes.data, handle

    def peer_open(self, handle):
        return _FakePeerCtx.registry[bytes(handle)].ctypes.data

    def peer_close(self, addr):
        pass

    def peer_free(self, addr):
        pass

    def peer_signal(self, flag_addr, value, stream=None):
        import ctypes
        ctypes.c_uint64.from_address(flag_addr).value = value

    def peer_wait(self, flags_addr, n, stride, value, timeout_s=10.0, stream=None):
        import ctypes, time
        t0 = time.time()
        for i in range(n):
            while ctypes.c_uint64.from_address(flags_addr + 8 * stride * i).value < value:
                if time.time() - t0 > timeout_s:
                    raise TimeoutError("peer_wait")
                time.sleep(0)


def test_peer_gather_protocol_threads():
    """Four ranks x 40 steps with random stalls: the root always reads the scores of the step it collected from every
    rank (no buffer is overwritten before it was released), uneven shards, two buffers."""
    import ctypes, threading, time, random
    from openwakeword_b200.distributed import PeerGather, shard_range
    world, n_total, n_cols, steps = 4, 10, 3, 40
    barrier = threading.Barrier(world)
    box, errors, seen = [None] * world, [], []

    def expected(k, row):
        return np.arange(n_cols, dtype=np.float32) + 100.0 * k + row

    def run(rank):
        try:
            rnd = random.Random(rank)

            def exchange(obj):
                box[rank] = obj
                barrier.wait()
                out = list(box)
                barrier.wait()
                return out
            pg = PeerGather(_FakePeerCtx(), n_total, n_cols, rank, world, exchange, slots=2, timeout_s=20.0)
            lo, hi = shard_range(n_total, rank, world)
            for k in range(1, steps + 1):
                if rank == pg.root and k > 1:
                    pg.release(k - 1)
                pg.begin(k)
                time.sleep(rnd.random() * 0.003)
                dst = np.ctypeslib.as_array(ctypes.cast(pg.dest(k), ctypes.POINTER(ctypes.c_float)), ((hi - lo), n_cols))
                for r in range(lo, hi):
                    dst[r - lo] = expected(k, r)
                pg.publish(k)
                if rank == pg.root:
                    addr = pg.collect(k)
                    time.sleep(rnd.random() * 0.003)              # a slow consumer
                    full = np.ctypeslib.as_array(ctypes.cast(addr, ctypes.POINTER(ctypes.c_float)), (n_total, n_cols)).copy()
                    seen.append((k, full))
            pg.close()
        except Exception as e:                                    # noqa: BLE001
            errors.append((rank, repr(e)))
            try:
                barrier.abort()
            except Exception:
                pass

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert not errors, errors
    assert len(seen) == steps
    for k, full in seen:
        want = np.stack([expected(k, r) for r in range(n_total)])
        np.testing.assert_array_equal(full, want)


def test_peer_gather_rejects_single_buffer():
    from openwakeword_b200.distributed import PeerGather
    with pytest.raises(ValueError):
        PeerGather(_FakePeerCtx(), 4, 1, 0, 1, lambda o: [o], slots=1)
