"""Host-side logic of the multi-GPU path, on CPU: nnz-balanced row partition, evaluation shards, and
the handle exchange / scalar reduction plumbing over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_partition_rows_balanced_and_aligned():
    from idgrec.dist import partition_rows, shard_range
    rng = np.random.default_rng(0)
    deg = np.minimum(rng.zipf(1.5, 20000), 5000)
    indptr = np.concatenate([[0], np.cumsum(deg)])
    for world in (1, 2, 4, 8):
        b = partition_rows(indptr, world)
        assert b[0] == 0 and b[-1] == 20000 and len(b) == world + 1
        assert all(b[i] <= b[i + 1] for i in range(world)) and all(x % 128 == 0 for x in b[:-1])
        work = [indptr[b[i + 1]] - indptr[b[i]] + 8 * (b[i + 1] - b[i]) for i in range(world)]
        assert max(work) <= 1.15 * (sum(work) / world) + 128 * 5000
    # degenerate: fewer rows than ranks
    b = partition_rows(np.array([0, 3, 5]), 4)
    assert b[0] == 0 and b[-1] == 2 and all(b[i] <= b[i + 1] for i in range(4))
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_range(2, 3, 4) == (2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the exchange PeerSlab does with CUDA-IPC handles: 64 opaque bytes per rank, same order everywhere
        handles = [None] * world
        dist.all_gather_object(handles, bytes([rank]) * 64)
        ok = all(h == bytes([r]) * 64 for r, h in enumerate(handles))
        # the reduction Test() does: per-shard float64 metric sums added across ranks
        from idgrec.dist import shard_range
        vals = np.arange(101, dtype=np.float64)
        s, e = shard_range(len(vals), rank, world)
        t = torch.tensor([vals[s:e].sum(), float(e - s)], dtype=torch.float64)
        dist.all_reduce(t)
        q.put((rank, ok, t.tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange_and_reduce():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    for rank, ok, t in out:
        assert ok and t == [5050.0, 101.0]


def _sample_check_worker(rank, world, port, q, same):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import utility.utility_train.trainer as trainer
        E = 1000
        rng = np.random.default_rng(7 if same else 7 + rank)      # equal seeds <=> equal negatives + permutation
        t = torch.from_numpy(np.stack([rng.integers(0, 500, E), rng.permutation(E)]).astype(np.int64))
        try:
            trainer._check_same_samples(t, E)
            q.put((rank, "ok"))
        except RuntimeError as e:
            q.put((rank, "raised: " + str(e)[:40]))
    finally:
        dist.destroy_process_group()


def _run2(target, *extra):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=target, args=(r, 2, port, q) + extra) for r in range(2)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    return out


def test_gloo_world2_rank_sample_checksum():
    """Row-partitioned training evaluates the loss redundantly on every rank (no gradient reduction), so all ranks must draw the
    same negatives and the same shuffle: trainer._check_same_samples passes on equal draws and raises on every rank otherwise."""
    assert [r for _, r in _run2(_sample_check_worker, True)] == ["ok", "ok"]
    out = _run2(_sample_check_worker, False)
    assert all(r.startswith("raised: ranks drew different epoch samples") for _, r in out), out
