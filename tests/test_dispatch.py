"""Host logic of the multi-GPU path on CPU: gloo, world_size 2 (no GPU needed)."""
import os
import socket

import pytest
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from larynx_b200 import dispatch


def test_sharding_is_a_partition():
    for n, w in ((10, 3), (7, 8), (0, 2), (16, 4)):
        seen = sorted(i for r in range(w) for i in dispatch.shard_round_robin(n, r, w))
        assert seen == list(range(n))
    costs = [5, 1, 9, 3, 3, 7, 2, 8]
    parts = dispatch.shard_by_cost(costs, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(costs)))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def make():
            calls.append(1)
            return (np.arange(100_003) % 251).astype(np.uint8)

        blob = dispatch.broadcast_blob(make if rank == 0 else None)
        assert blob.dtype == torch.uint8 and blob.numel() == 100_003
        assert int(blob.to(torch.int64).sum()) == int(((np.arange(100_003) % 251)).sum())
        assert len(calls) == (1 if rank == 0 else 0), "only the source rank packs the voice"
        mine = dispatch.shard_round_robin(9, rank, world)
        local = {i: ("utt%d" % i, rank) for i in mine}
        ordered = dispatch.gather_in_order(local, 9)
        assert [o[0] for o in ordered] == ["utt%d" % i for i in range(9)]
        assert [o[1] for o in ordered] == [i % world for i in range(9)]
        out.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_bucket_by_length_covers_every_sentence_once():
    rs = np.random.RandomState(0)
    lens = rs.randint(1, 300, 97)
    for batch, cap in ((8, None), (16, 2000), (1, None)):
        buckets = dispatch.bucket_by_length(lens, batch, cap)
        flat = sorted(i for b in buckets for i in b)
        assert flat == list(range(97))
        assert all(1 <= len(b) <= batch for b in buckets)
        if cap is not None:
            assert all(len(b) == 1 or max(lens[i] for i in b) * len(b) <= cap for b in buckets)
        # similar lengths end up together: within a bucket the spread is small compared with the whole set
        if batch > 1:
            spread = np.mean([max(lens[i] for i in b) - min(lens[i] for i in b) for b in buckets])
            assert spread < (lens.max() - lens.min()) / 4


def test_batch_scheduler_orders_results_and_propagates_errors():
    from larynx_b200.scheduler import BatchScheduler

    rs = np.random.RandomState(1)
    rows = [rs.randint(1, 46, n) for n in rs.randint(0, 120, 41)]
    rows[7] = np.zeros(0, dtype=np.int64)
    calls = []

    def fake(batch_rows, seed):
        calls.append([len(r) for r in batch_rows])
        return [np.full(len(r) * 3, len(r), dtype=np.int16) for r in batch_rows]

    sch = BatchScheduler(max_batch=8, max_cost=400, workers=3, synth=fake)
    out = sch.synthesize(rows)
    assert [a.shape[0] for a in out] == [3 * len(r) for r in rows]
    assert all(len(c) <= 8 and max(c) * len(c) <= 400 or len(c) == 1 for c in calls)
    assert sorted(n for c in calls for n in c) == sorted(len(r) for r in rows if len(r))  # empty sentences never reach the device
    seen = [i for i, _ in sch.stream(rows)]
    assert seen == list(range(len(rows)))

    def broken(batch_rows, seed):
        raise ValueError("boom")

    import pytest

    with pytest.raises(ValueError):
        BatchScheduler(max_batch=4, workers=2, synth=broken).synthesize(rows)
    with pytest.raises(ValueError):
        list(BatchScheduler(max_batch=4, workers=2, synth=broken).stream(rows))


def test_scheduler_reserve_runs_the_costliest_batch_once_on_every_worker():
    import threading

    from larynx_b200.scheduler import BatchScheduler

    seen = []
    lock = threading.Lock()

    def fake(rows, seed):
        with lock:
            seen.append((threading.get_ident(), sum(len(r) for r in rows), len(rows)))
        return [np.zeros(len(r), np.int16) for r in rows]

    sch = BatchScheduler(max_batch=4, max_cost=None, workers=3, synth=fake)
    try:
        rows = [np.ones(n, np.int64) for n in (3, 50, 7, 48, 2, 49, 51, 9, 1)]
        sch.reserve(rows)
        assert len(seen) == 3 and len({t for t, _, _ in seen}) == 3  # one job per worker thread
        assert all(total == 9 + 48 + 49 + 50 and n == 4 for _, total, n in seen)  # ... the costliest micro-batch (4 x 50 padded)
        out = sch.synthesize(rows)  # the scheduler still works afterwards
        assert [len(a) for a in out] == [len(r) for r in rows]
        sch.reserve([])  # nothing to plan: no-op

        def broken(rows, seed):
            raise ValueError("boom")

        bad = BatchScheduler(max_batch=2, workers=2, synth=broken)
        with pytest.raises(ValueError):
            bad.reserve(rows)
        bad.close()
    finally:
        sch.close()
