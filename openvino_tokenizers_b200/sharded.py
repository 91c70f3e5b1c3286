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


def allgather_ragged_slots(ids_buf: torch.Tensor, begins: torch.Tensor, ends: torch.Tensor, out_ids: torch.Tensor = None, group=None,
                           async_op: bool = False):
    """All-gatherv of ragged int32 rows without a host synchronisation or staging copy.

    Every rank contributes its whole fixed-capacity id buffer (ids_buf: int32[cap], rows at [begins[i], ends[i]), the
    tail beyond the last row unspecified) straight from where the tokenizer wrote it; rank r's buffer lands in slot r of
    the result and its row offsets are shifted by r * cap.  The gaps between the slots are legal in the reference's
    ragged representation (begins/ends need not be contiguous, e.g. after Truncate).  Requires the same cap and the same
    number of rows on every rank (the weak-scaling layout: equal shards).
    returns (begins int32[W*B], ends int32[W*B], ids int32[W*cap]), identical on every rank.  With async_op=True the
    collectives are only enqueued (the caller's stream does not wait for them) and a function is returned that waits and
    yields the tuple — this is how a multi-block step overlaps the exchange of block k with the tokenisation of block k+1.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cap, B = ids_buf.numel(), begins.numel()
    dev = ids_buf.device
    if out_ids is None:
        out_ids = torch.empty(world * cap, dtype=torch.int32, device=dev)
    w1 = dist.all_gather_into_tensor(out_ids, ids_buf, group=group, async_op=async_op)
    be = (torch.cat([begins, ends]) + rank * cap).to(torch.int32)  # [2 * B], offsets into the gathered buffer
    all_be = torch.empty(world * 2 * B, dtype=torch.int32, device=dev)
    w2 = dist.all_gather_into_tensor(all_be, be, group=group, async_op=async_op)

    def finish():
        if async_op:
            w1.wait()
            w2.wait()
        v = all_be.view(world, 2, B)
        return v[:, 0].reshape(-1), v[:, 1].reshape(-1), out_ids
    return finish if async_op else finish()


class PeerGather:
    """Emit fused with the all-gatherv over NVLink peer memory: the result buffers of every rank live in torch
    symmetric memory; `b200tok_split_bpe_run_sharded` stores this rank's compacted id rows into all of them from inside its
    compaction kernel, and one symmetric-memory barrier on the stream orders the ranks.  Equal shards (rows, capacity) per rank.
    Result layout = allgather_ragged_slots: rank r's rows in slot r, offsets shifted by r * cap."""

    def __init__(self, rows_per_rank: int, slot_capacity: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _capi as K
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PeerGather supports up to 8 ranks (one NVSwitch domain)")
        self.rows, self.cap = int(rows_per_rank), int(slot_capacity)
        mk = lambda n: symm.empty(n, dtype=torch.int32, device=device)
        self.ids, self.begins, self.ends = mk(self.world * self.cap), mk(self.world * self.rows), mk(self.world * self.rows)
        self._h = [symm.rendezvous(t, self.group) for t in (self.ids, self.begins, self.ends)]
        self.n = torch.zeros(1, dtype=torch.int64, device=device)
        po = K.PeerOut()
        po.world, po.rank, po.slot_capacity, po.rows_per_rank = self.world, self.rank, self.cap, self.rows
        for p in range(self.world):
            po.ids[p], po.begins[p], po.ends[p] = (int(h.buffer_ptrs[p]) for h in self._h)
        self._po = po

    def run(self, pipe, db):
        """pipe: runtime.TokenizerPipeline (kind 'bpe'); db: runtime.DeviceBatch of this rank's shard.  Asynchronous on the current
        torch stream up to the barrier; returns (begins, ends, ids) views of this rank's copy of the gathered result."""
        import ctypes as C
        from . import _capi as K
        rin = K.RaggedStrings(db.rb.data_ptr(), db.re.data_ptr(), db.n_rows, db.begins.data_ptr(), db.ends.data_ptr(), db.n_elems,
                              db.chars.data_ptr(), db.n_chars, None, K.MEM_DEVICE)
        self._h[0].barrier(channel=1)   # nobody still reads the previous result (readers are ordered before this on their streams)
        K.check(K.lib().b200tok_split_bpe_run_sharded(pipe.split1.handle, pipe.tok.handle, C.byref(rin), C.byref(self._po),
                                                      C.c_void_p(self.n.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self._h[0].barrier(channel=0)   # every rank's stores into my buffers are complete and visible after this
        return self.begins, self.ends, self.ids
