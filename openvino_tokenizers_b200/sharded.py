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
    Result layout = allgather_ragged_slots: rank r's rows in slot r, offsets shifted by r * cap.

    The result buffers are double-buffered (step k uses set k & 1), so ONE barrier per step is enough: a rank overwrites set k & 1
    again in step k + 2, after it has passed the barrier of step k + 1 — which every peer reaches only when everything it enqueued
    before it, readers of its step-k result included, has finished.  `begins / ends / ids` are the set of the latest step."""

    def __init__(self, rows_per_rank: int, slot_capacity: int, device, group=None, wire16: bool = False, multicast: bool = False,
                 double_buffer: bool = True):
        """wire16: ship ids over NVLink as u16 (every id < 65 535) into symmetric staging buffers and widen them locally after the
        barrier — halves the NVLink bytes, worth a few percent from about four ranks on.  multicast: use NVLS multimem.st through the
        switch instead of one store per peer (measured slower than wide unicast stores for this access pattern on B200: off by default)."""
        import torch.distributed._symmetric_memory as symm
        from . import _capi as K
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PeerGather supports up to 8 ranks (one NVSwitch domain)")
        self.rows, self.cap, self.wire16, self.device = int(rows_per_rank), int(slot_capacity), bool(wire16), device
        self.n = torch.zeros(1, dtype=torch.int64, device=device)
        self.multicast = False
        self.step = 0
        self._sets = []
        mk = lambda n, dt=torch.int32: symm.empty(n, dtype=dt, device=device)
        for _ in range(2 if (double_buffer and self.world > 1) else 1):
            begins, ends = mk(self.world * self.rows), mk(self.world * self.rows)
            if self.wire16:
                ids = torch.empty(self.world * self.cap, dtype=torch.int32, device=device)       # local, filled by the widening pass
                ids16 = mk(self.world * self.cap, torch.int16)
                syms = (ids16, begins, ends)
            else:
                ids = mk(self.world * self.cap)
                ids16 = None
                syms = (ids, begins, ends)
            h = [symm.rendezvous(t, self.group) for t in syms]
            po = K.PeerOut()
            po.world, po.rank, po.slot_capacity, po.rows_per_rank, po.wire16 = self.world, self.rank, self.cap, self.rows, int(self.wire16)
            for p in range(self.world):
                if self.wire16:
                    po.ids16[p] = int(h[0].buffer_ptrs[p])
                    po.ids[p] = ids.data_ptr() if p == self.rank else None
                else:
                    po.ids[p] = int(h[0].buffer_ptrs[p])
                po.begins[p], po.ends[p] = int(h[1].buffer_ptrs[p]), int(h[2].buffer_ptrs[p])
            if multicast and not self.wire16 and self.world > 1:      # NVLS: one store through the switch reaches every rank's copy
                try:
                    mc = [int(x.multicast_ptr) for x in h]
                    if all(mc):
                        po.ids_mc, po.begins_mc, po.ends_mc = mc
                        self.multicast = True
                except Exception:
                    pass
            self._sets.append(dict(begins=begins, ends=ends, ids=ids, ids16=ids16, h=h, po=po))
        self._cur = self._sets[0]

    begins = property(lambda self: self._cur["begins"])
    ends = property(lambda self: self._cur["ends"])
    ids = property(lambda self: self._cur["ids"])

    def run(self, pipe, db):
        """pipe: runtime.TokenizerPipeline (BPE or WordPiece); db: runtime.DeviceBatch of this rank's shard.  Asynchronous on the current
        torch stream up to the barrier; returns (begins, ends, ids) views of this rank's copy of the gathered result."""
        import ctypes as C
        from . import _capi as K
        cur = self._cur = self._sets[self.step % len(self._sets)]
        self.step += 1
        rin = K.RaggedStrings(db.rb.data_ptr(), db.re.data_ptr(), db.n_rows, db.begins.data_ptr(), db.ends.data_ptr(), db.n_elems,
                              db.chars.data_ptr(), db.n_chars, None, K.MEM_DEVICE)
        if len(self._sets) == 1:
            cur["h"][0].barrier(channel=1)   # nobody still reads the previous result (readers are ordered before this on their streams)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if pipe.kind == "bpe":
            K.check(K.lib().b200tok_split_bpe_run_sharded(pipe.split1.handle, pipe.tok.handle, C.byref(rin), C.byref(cur["po"]),
                                                          C.c_void_p(self.n.data_ptr()), st))
        else:
            K.check(K.lib().b200tok_split_wordpiece_run_sharded(pipe.split1.handle, pipe.split2.handle, pipe.tok.handle, C.byref(rin),
                                                                C.c_int32(pipe.unk), C.byref(cur["po"]), C.c_void_p(self.n.data_ptr()), st))
        cur["h"][0].barrier(channel=0)   # every rank's stores into my buffers are complete and visible after this
        if self.wire16:
            dev_index = self.device.index if isinstance(self.device, torch.device) else int(self.device)
            K.check(K.lib().b200tok_peer_expand_run(int(dev_index or 0), C.byref(cur["po"]), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return cur["begins"], cur["ends"], cur["ids"]


class PullGather:
    """All-gatherv by PULL over NVLink peer memory.  Every rank tokenises its shard with the ordinary one-GPU call —
    at its one-GPU speed, no remote stores inside the tokenizer kernel — into peer-mapped (torch symmetric memory) source buffers:
    compact begins / ends / id count, and the ids packed to 16 bits when every id fits (`b200tok_peer_pack_run`), else the i32 ids
    themselves.  After ONE symmetric-memory barrier on the stream, `b200tok_peer_pull_run` reads every peer's source buffers with
    16-byte loads over NVLink and widens them straight into this rank's i32 result (transfer and widening are one pass).  The source
    buffers are double-buffered, so the next step may overwrite them without a second barrier: a peer that still reads step k's
    buffers has not yet arrived at the barrier of step k + 1, which this rank must pass before step k + 2 writes them again.
    Result layout = allgather_ragged_slots / PeerGather: rank r's rows in slot r, offsets shifted by r * cap.  Equal shards per rank."""

    def __init__(self, rows_per_rank: int, slot_capacity: int, device, group=None, wire16: bool = False):
        import torch.distributed._symmetric_memory as symm
        from . import _capi as K
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PullGather supports up to 8 ranks (one NVSwitch domain)")
        self.rows, self.cap, self.wire16, self.device = int(rows_per_rank), (int(slot_capacity) + 7) & ~7, bool(wire16), device
        self.dev_index = int((device.index if isinstance(device, torch.device) else device) or 0)
        W, R, cap = self.world, self.rows, self.cap
        self._stride = cap + 8                                     # ids per parity of the source buffer (padded: sources are read in groups of 8)
        self._src_be = symm.empty(4 * R, dtype=torch.int32, device=device)          # [parity][begins | ends][rows]
        self._src_tot = symm.empty(2, dtype=torch.int64, device=device)             # [parity]
        self._src_ids = symm.empty(2 * self._stride, dtype=torch.int16 if self.wire16 else torch.int32, device=device)
        self._src_tot.zero_()
        self._h = [symm.rendezvous(t, self.group) for t in (self._src_be, self._src_tot, self._src_ids)]
        self.ids = torch.empty(W * cap + 8, dtype=torch.int32, device=device)       # the gathered result (local memory only)
        self.begins = torch.empty(W * R, dtype=torch.int32, device=device)
        self.ends = torch.empty(W * R, dtype=torch.int32, device=device)
        self.step = 0
        self.multicast = False
        self.compact_slots = True          # every slot holds its rank's rows back to back
        self._pull = []
        esz = 2 if self.wire16 else 4
        for par in (0, 1):
            q = K.PeerPull()
            q.world, q.rank, q.wire16, q.skip_self_ids = W, self.rank, int(self.wire16), int(self.wire16)
            q.slot_capacity, q.rows_per_rank = cap, R
            q.ids, q.begins, q.ends = self.ids.data_ptr(), self.begins.data_ptr(), self.ends.data_ptr()
            for p in range(W):
                be, tot, ids = (int(h.buffer_ptrs[p]) for h in self._h)
                q.src_begins[p] = be + 4 * (2 * par) * R
                q.src_ends[p] = be + 4 * (2 * par + 1) * R
                q.src_total[p] = tot + 8 * par
                if self.wire16:
                    q.src_ids16[p] = ids + esz * par * self._stride
                else:
                    q.src_ids[p] = ids + esz * par * self._stride
            self._pull.append(q)

    @property
    def n(self):
        """This rank's id count of the last step (device int64[1])."""
        return self._src_tot[(self.step - 1) & 1:((self.step - 1) & 1) + 1]

    def run(self, pipe, db):
        """pipe: runtime.TokenizerPipeline; db: runtime.DeviceBatch of this rank's shard (db.n_rows == rows_per_rank).  Asynchronous on the
        current torch stream; returns (begins, ends, ids) of the gathered result held by this rank."""
        import ctypes as C
        from . import _capi as K
        if db.n_rows != self.rows:
            raise ValueError("PullGather: every rank passes rows_per_rank rows")
        par = self.step & 1
        self.step += 1
        R, cap = self.rows, self.cap
        L = K.lib()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rin = K.RaggedStrings(db.rb.data_ptr(), db.re.data_ptr(), db.n_rows, db.begins.data_ptr(), db.ends.data_ptr(), db.n_elems,
                              db.chars.data_ptr(), db.n_chars, None, K.MEM_DEVICE)
        be, tot = self._src_be.data_ptr(), self._src_tot.data_ptr() + 8 * par
        if self.wire16:      # the i32 ids go straight to their final place (this rank's slot of its own result); peers read the packed copy
            ids_ptr = self.ids.data_ptr() + 4 * self.rank * cap
        else:
            ids_ptr = self._src_ids.data_ptr() + 4 * par * self._stride
        out = K.RaggedIds(be + 4 * (2 * par) * R, be + 4 * (2 * par + 1) * R, ids_ptr, cap, 0, tot, K.MEM_DEVICE)
        pipe._call(rin, out, st)
        if self.wire16:
            K.check(L.b200tok_peer_pack_run(self.dev_index, C.c_void_p(ids_ptr), C.c_void_p(tot), C.c_int64(cap),
                                            C.c_void_p(self._src_ids.data_ptr() + 2 * par * self._stride), st))
        self._h[0].barrier(channel=0)       # every rank's source buffers of this step are complete and visible after this
        K.check(L.b200tok_peer_pull_run(self.dev_index, C.byref(self._pull[par]), st))
        return self.begins, self.ends, self.ids
