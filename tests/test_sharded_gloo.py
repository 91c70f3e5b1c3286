"""CPU tier: the N>1 host logic (row sharding + all-gatherv of ragged rows) on world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openvino_tokenizers_b200.sharded import allgather_ragged_slots, allgatherv_ragged, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)
    counts_all = rng.integers(0, 9, size=n_rows).astype(np.int32)
    ids_all = rng.integers(0, 50000, size=int(counts_all.sum())).astype(np.int32)
    ends = np.cumsum(counts_all)
    lo, hi = shard_bounds(n_rows, world, rank)
    t0 = int(ends[lo - 1]) if lo > 0 else 0
    t1 = int(ends[hi - 1]) if hi > 0 else 0
    b, e, ids = allgatherv_ragged(torch.from_numpy(ids_all[t0:t1].copy()), torch.from_numpy(counts_all[lo:hi].copy()))
    ok = (np.array_equal(ids.numpy(), ids_all) and np.array_equal(e.numpy(), ends.astype(np.int32))
          and np.array_equal(b.numpy(), (ends - counts_all).astype(np.int32)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [1, 7, 1000])
def test_allgatherv_ragged_world2(n_rows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _slots_worker(rank, world, port, n_rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cap = 10 * n_rows + 5
    rows = []
    for r in range(world):                       # every rank can rebuild every shard: seeded per shard
        rng = np.random.default_rng(100 + r)
        counts = rng.integers(0, 9, size=n_rows).astype(np.int32)
        ids = rng.integers(0, 50000, size=int(counts.sum())).astype(np.int32)
        ends = np.cumsum(counts).astype(np.int32)
        rows.append((ends - counts, ends, ids))
    b, e, ids = rows[rank]
    buf = torch.full((cap,), -1, dtype=torch.int32)
    buf[: len(ids)] = torch.from_numpy(ids)
    gb, ge, gids = allgather_ragged_slots(buf, torch.from_numpy(b.copy()), torch.from_numpy(e.copy()))
    ok = gb.numel() == world * n_rows and gids.numel() == world * cap
    for r in range(world):
        rb, re_, rids = rows[r]
        for i in range(n_rows):
            k = r * n_rows + i
            ok &= bool(np.array_equal(gids[int(gb[k]): int(ge[k])].numpy(), rids[rb[i]: re_[i]]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [1, 64])
def test_allgather_ragged_slots_world2(n_rows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slots_worker, args=(r, 2, port, n_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 65536, 262144):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
