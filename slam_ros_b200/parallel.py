"""Host-side helpers for the two multi-GPU modes (one process per GPU).

* Monte-Carlo batches: filters are independent -> filter i lives on rank i % world; no collective on
  the data path.  Results are gathered on the host only for parity checks.
* Row-sharded covariance: 64-row tile rows are dealt round-robin (rb % world), which balances the
  upper-triangle work (row r has n - r elements).  The only exchange per matched line is the pair of
  H-column slices (done inside libekfcuda with NCCL).

Nothing here touches a GPU; it is exercised on CPU with the gloo backend in tests/.
"""
import numpy as np

TILE = 64


def filters_of_rank(n_filters, rank, world):
    """Global ids of the filters rank `rank` owns (round-robin)."""
    return np.arange(rank, n_filters, world, dtype=np.int64)


def owner_of_row(r, world):
    return (int(r) // TILE) % world


def local_row(r, world):
    return ((int(r) // TILE) // world) * TILE + int(r) % TILE


def tiles_of_rank(n_live, rank, world):
    """Number of 64x64 upper-triangle tiles rank `rank` sweeps for a live dimension n_live."""
    T = (n_live + TILE - 1) // TILE
    return sum(T - rb for rb in range(rank, T, world))


def gather_filter_results(local_ids, local_values, n_filters, group=None):
    """All-gather per-filter result rows (e.g. poses) into global filter order.  Works on any backend."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    vals = torch.as_tensor(np.asarray(local_values, dtype=np.float64))
    width = vals.shape[1] if vals.ndim > 1 else 1
    vals = vals.reshape(len(local_ids), width)
    out = torch.zeros((n_filters, width), dtype=torch.float64)
    out[torch.as_tensor(np.asarray(local_ids))] = vals
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)   # disjoint rows: the sum is a gather
    return out.numpy()


def connect_shards(filt, device, group=None):
    """Row-sharded filter: all-gather the ranks' CUDA IPC handles with torch.distributed and map the peers'
    exchange buffers (ekf_shard_connect); if and only if every rank succeeded, every rank switches to the fused
    exchange (ekf_shard_use_fused).  Returns the common outcome."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor(list(filt.shard_ipc_handle()), dtype=torch.uint8, device=device)
    allh = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    ok = filt.shard_connect([bytes(h.cpu().tolist()) for h in allh])
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    all_ok = bool(int(flag[0]))
    if all_ok:                                  # two-phase: only switch paths when EVERY rank mapped its peers
        filt.shard_use_fused(True)
    return all_ok


def max_over_ranks(value, group=None):
    """Timing rule: a multi-GPU number is the max over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
