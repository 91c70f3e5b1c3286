"""2..8-GPU check of the two peer-memory exchanges — emit fused with peer stores (sharded.PeerGather) and all-gatherv by pull
(sharded.PullGather) — against the NCCL slot all-gather:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/peer_gather_check.py
Every rank tokenises its own C1 shard; both exchanges must give identical (begins, ends, ids-in-rows) on every rank."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cases
from openvino_tokenizers_b200 import runtime as R
from openvino_tokenizers_b200.sharded import PeerGather, PullGather, allgather_ragged_slots

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pipe = R.TokenizerPipeline("bpe", "gpt2_synth", device=local)
batch = cases.random_ascii_batch(rows, 512, 1234 + rank)
db = R.to_device(batch, dev)
cap = db.n_chars
wire16 = os.environ.get("B200TOK_WIRE16", "0") == "1"
pg = PeerGather(rows, cap, dev, wire16=wire16, multicast=os.environ.get("B200TOK_MULTICAST", "0") == "1")
o = pipe.run_device(db)
ref = allgather_ragged_slots(o["ids"][:cap], o["begins"], o["ends"])
b, e, ids = pg.run(pipe, db)
torch.cuda.synchronize()
rb, re_, rids = (x.cpu().numpy() for x in ref)
gb, ge, gids = b.cpu().numpy(), e.cpu().numpy(), ids.cpu().numpy()
ok = bool(np.array_equal(ge - gb, re_ - rb))            # same row lengths (the fused emit leaves rows at their worst-case positions)
tot = int((ge - gb).sum())
# every row of every rank: same ids (vectorised: gather both layouts into compact form)
def compact(bb, ee, xx):
    lens = (ee - bb).astype(np.int64)
    idx = np.repeat(bb.astype(np.int64) - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(int(lens.sum()))
    return xx[idx]
ok &= bool(np.array_equal(compact(gb, ge, gids), compact(rb, re_, rids)))
ok &= bool(int(pg.n.item()) == int((ge - gb)[rank * rows:(rank + 1) * rows].sum()))
# the pull exchange, twice (both parities of its double-buffered source buffers), 16-bit and 32-bit wire
pulls = {}
for w16 in (True, False):
    pl = PullGather(rows, cap, dev, wire16=w16)
    for _ in range(3):
        pb, pe, pids = pl.run(pipe, db)
        torch.cuda.synchronize()
        pb_, pe_, pids_ = pb.cpu().numpy(), pe.cpu().numpy(), pids.cpu().numpy()
        ok_p = bool(np.array_equal(pe_ - pb_, re_ - rb)) and bool(np.array_equal(compact(pb_, pe_, pids_), compact(rb, re_, rids)))
        ok_p &= bool(np.array_equal(pb_ // pl.cap, np.repeat(np.arange(world), rows)[: len(pb_)]) or tot == 0)      # every row inside its rank's slot
        ok_p &= int(pl.n.item()) == int((pe_ - pb_)[rank * rows:(rank + 1) * rows].sum())
        ok &= ok_p
    pulls[w16] = pl


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3


g = torch.empty(world * cap, dtype=torch.int32, device=dev)
t_nccl = timed(lambda: allgather_ragged_slots(pipe.run_device(db)["ids"][:cap], o["begins"], o["ends"], g))
t_peer = timed(lambda: pg.run(pipe, db))
t_local = timed(lambda: pipe.run_device(db))
t_pull16 = timed(lambda: pulls[True].run(pipe, db))
t_pull32 = timed(lambda: pulls[False].run(pipe, db))
flags = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    mb = rows * 512 * world / 1e6
    print(f"[wire16={wire16} multicast={pg.multicast}] peer-store gather and pull gather == nccl slot gather on every rank: {bool(flags.item())}; ids {tot}; ms/step local-only {t_local:.3f}, "
          f"tokenise+NCCL all-gather {t_nccl:.3f} ({mb / t_nccl * 1e3 / 1e3:.1f} GB/s text), tokenise+peer-store emit {t_peer:.3f} ({mb / t_peer:.1f} GB/s text), "
          f"tokenise+pull (16-bit wire) {t_pull16:.3f} ({mb / t_pull16:.1f} GB/s text), tokenise+pull (32-bit wire) {t_pull32:.3f}")
dist.destroy_process_group()
