"""Row sharding across GPUs and the all-gatherv of ragged token rows.

Rows (strings) are independent in every hot-path op, so a batch is cut into contiguous row blocks, one per rank
(one process per GPU); each rank tokenises its block with the CUDA path; the only exchange is one all-gatherv of
the ragged id rows: (1) all-gather of the per-rank id counts and per-row token counts, (2) one `ncclAllGather`
of the id payload padded to the largest shard, (3) a local compaction + exclusive scan that rebuilds the global
`(begins, ends, ids)` on every rank.  Works on any torch.distributed backend (NCCL on the GPUs; gloo in the CPU
tests of this logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int):
    """Contiguous row block [lo, hi) of `rank` (docs are fixed-length in the benchmark configs, so bytes balance)."""
    return n_rows * rank // world, n_rows * (rank + 1) // world


def allgatherv_ragged(local_ids: torch.Tensor, local_row_counts: torch.Tensor, group=None):
    """All-gatherv of ragged int32 rows.

    local_ids:        int32[T_r]   this rank's token ids, rows concatenated
    local_row_counts: int32[B_r]   tokens per local row
    returns (begins int32[B], ends int32[B], ids int32[T]) for the whole batch, identical on every rank.
    """
    world = dist.get_world_size(group)
    dev = local_ids.device
    meta = torch.tensor([local_ids.numel(), local_row_counts.numel()], dtype=torch.int64, device=dev)
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.view(world, 2).cpu()          # the collective sizes are host arguments: one small sync
    t_sizes, b_sizes = metas[:, 0].tolist(), metas[:, 1].tolist()
    t_max, b_max = max(max(t_sizes), 1), max(max(b_sizes), 1)

    pad_ids = torch.zeros(t_max, dtype=torch.int32, device=dev)
    pad_ids[: local_ids.numel()] = local_ids
    all_ids = torch.empty(world * t_max, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_ids, pad_ids, group=group)

    pad_cnt = torch.zeros(b_max, dtype=torch.int32, device=dev)
    pad_cnt[: local_row_counts.numel()] = local_row_counts
    all_cnt = torch.empty(world * b_max, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_cnt, pad_cnt, group=group)

    ids = torch.cat([all_ids[r * t_max: r * t_max + t_sizes[r]] for r in range(world)])
    counts = torch.cat([all_cnt[r * b_max: r * b_max + b_sizes[r]] for r in range(world)])
    ends = torch.cumsum(counts, 0, dtype=torch.int64).to(torch.int32)
    begins = ends - counts
    return begins, ends, ids
