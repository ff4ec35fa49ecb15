"""N > 1 host logic on CPU: world_size-2 gloo process group (rendezvous on 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from slam_ros_b200 import parallel as par


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_filters, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = par.filters_of_rank(n_filters, rank, world)
        # each rank "runs" its own filters: result row = f(global id), a stand-in for the per-filter pose
        vals = np.stack([ids * 1.5, ids * -2.0, ids % 7], axis=1).astype(np.float64)
        full = par.gather_filter_results(ids, vals, n_filters)
        t = par.max_over_ranks(10.0 + rank)
        q.put((rank, full, t, len(ids)))
    finally:
        dist.destroy_process_group()


def test_monte_carlo_sharding_world2():
    world, n_filters = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_filters, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.arange(n_filters)
    want = np.stack([g * 1.5, g * -2.0, g % 7], axis=1).astype(np.float64)
    counts = 0
    for rank, full, t, cnt in res:
        assert np.array_equal(full, want)
        assert t == 11.0                      # max over ranks
        counts += cnt
    assert counts == n_filters


def test_filter_partition_is_disjoint_and_complete():
    for world in (1, 2, 4, 8):
        allids = np.concatenate([par.filters_of_rank(4096, r, world) for r in range(world)])
        assert np.array_equal(np.sort(allids), np.arange(4096))
        sizes = [len(par.filters_of_rank(4096, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_row_block_ownership_balances_the_triangle():
    n = 80003
    for world in (2, 4, 8):
        tiles = [par.tiles_of_rank(n, r, world) for r in range(world)]
        T = (n + 63) // 64
        assert sum(tiles) == T * (T + 1) // 2
        assert (max(tiles) - min(tiles)) / max(tiles) < 0.02      # cyclic deal: < 2 % imbalance
    # local row mapping is a bijection onto [0, rows_local)
    world = 4
    for rank in range(world):
        rows = [r for r in range(0, 64 * 40) if par.owner_of_row(r, world) == rank]
        loc = sorted(par.local_row(r, world) for r in rows)
        assert loc == list(range(len(rows)))


class _StubShard:
    """Stands in for a row-sharded EkfFilter: records the handles it is asked to connect."""

    def __init__(self, rank, fail=False):
        self.rank = rank; self.fail = fail; self.seen = None; self.fused = False

    def shard_ipc_handle(self):
        return bytes([self.rank]) * 64

    def shard_connect(self, handles):
        self.seen = [bytes(h) for h in handles]
        return not self.fail

    def shard_use_fused(self, on=True):
        self.fused = bool(on)


def _connect_worker(rank, world, port, fail_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stub = _StubShard(rank, fail=(rank == fail_rank))
        ok = par.connect_shards(stub, "cpu")
        q.put((rank, ok, stub.seen, stub.fused))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_shard_handle_exchange_world2(fail_rank):
    """connect_shards: every rank receives every rank's 64-byte IPC handle in rank order, and the ranks agree on
    the outcome (one rank failing to map its peers keeps ALL ranks on the NCCL exchange)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_connect_worker, args=(r, world, port, fail_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, seen, fused in res:
        assert seen == [bytes([r]) * 64 for r in range(world)]
        assert ok == (fail_rank < 0) and fused == ok        # no rank switches paths unless all mapped their peers
