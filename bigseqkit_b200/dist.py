"""Multi-GPU plumbing: one process per GPU, shards partition one per rank.

The map-only operators (seq / subseq / translate / grep / locate) need no data-path collective: every rank
runs its record-aligned shard and `output_offsets` (one all-gather of sizes) tells it where its bytes go in the
merged output -- the job the reference does with an MPI token ring in FileStore
(bigseqkit-lib/helper.go:399-431).  Two operators have a real exchange step:

  stats   StatsReduce (bigseqkit-lib/stats.go:128-137, bigseqkit/stats.go:91)   -> one all-reduce / all-gather
  rmdup   GroupByKey  (bigseqkit/rmdup.go:97)                                   -> one all-gather of 16-byte
          fingerprints; the first occurrence in GLOBAL input order survives

`torch.distributed` is only plumbing here (NCCL over NVLink on GPUs, gloo in the CPU tests); the arithmetic is in
libbsk.so.
"""
import numpy as np
import torch
import torch.distributed as dist

from .api import Operator


# ---------------------------------------------------------------------------- shard planner
def _is_record_start(buf, pos, fastq):
    """`pos` starts a record: FASTA '\\n>' ; FASTQ '\\n@' unless preceded by '\\n+' (SURVEY C.1)."""
    if pos == 0:
        return True
    marker = 0x40 if fastq else 0x3E
    if buf[pos - 1] != 0x0A or buf[pos] != marker:
        return False
    if fastq and pos >= 3 and buf[pos - 3] == 0x0A and buf[pos - 2] == 0x2B:
        return False
    return True


def shard_bounds(buf, world):
    """Cut `buf` (bytes-like) into `world` contiguous record-aligned byte ranges: every cut moves forward to the
    next record start (drv helper.go:148-178 leaves this to IgnisHPC's PlainFile).  Returns world+1 offsets."""
    mv = memoryview(buf).cast("B")
    n = len(mv)
    fastq = n > 0 and mv[0] == 0x40
    cuts = [0]
    for r in range(1, world):
        pos = max(cuts[-1], n * r // world)
        while pos < n and not _is_record_start(mv, pos, fastq):
            pos += 1
        cuts.append(pos)
    cuts.append(n)
    return cuts


# ---------------------------------------------------------------------------- collectives
def _world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def output_offsets(n_local, device="cpu", group=None):
    """Global byte offset of this rank's output and the total size: all-gather of one int64 per rank."""
    rank, world = _world(group)
    if world == 1:
        return 0, int(n_local)
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    sizes = [int(s.item()) for s in sizes]
    return sum(sizes[:rank]), sum(sizes)


def stats_allreduce(op, device="cpu", nbins=65536, group=None):
    """Merge the Stats totals of all ranks into every rank's operator (sum semantics).

    Lengths below `nbins` travel as a dense uint64 histogram through ONE all-reduce (bsk_stats_dense_device fills
    it on the device); the few scalars and any longer lengths go through an all-gather of python objects."""
    rank, world = _world(group)
    if world == 1:
        return op
    res = op.stats_result()
    hist = torch.zeros(nbins, dtype=torch.int64, device=device)
    op.stats_dense_device(hist.data_ptr(), nbins)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    tail = {"long": [(l, c) for l, c in res["hist"] if l >= nbins], "q20": res["q20"], "q30": res["q30"],
            "gap": res["sum_gap"], "type": res["type"], "num": res["num"]}
    tails = [None] * world
    dist.all_gather_object(tails, tail, group=group)
    # rebuild the operator's totals from the reduced values
    op.reset()
    h = hist.cpu().numpy()
    nz = np.nonzero(h)[0]
    pairs = [(int(l), int(h[l])) for l in nz]
    for t in tails:
        pairs.extend(t["long"])
    merged = {}
    for l, c in pairs:
        merged[l] = merged.get(l, 0) + c
    # the type column comes from the first rank that saw a record (reference: partition 0's first record)
    typ = next((t["type"] for t in tails if t["num"] > 0), "")
    op.stats_add(sorted(merged.items()), q20=sum(t["q20"] for t in tails), q30=sum(t["q30"] for t in tails),
                 sum_gap=sum(t["gap"] for t in tails), type=typ)
    return op


def rmdup_union(op, d_in_ptr, nbytes, device="cpu", group=None):
    """rmdup over all ranks' shards (rank order == input order): hash locally, all-gather the 16-byte fingerprints,
    drop every local record whose fingerprint occurs earlier in global order.  Returns the bsk_out of the survivors
    (device pointers) and the number of local records."""
    rank, world = _world(group)
    cap = max(1, nbytes // 2 + 1)
    if world == 1:
        out = op.call_device(d_in_ptr, nbytes)
        return out, int(out.n_records)
    # upper bound on the records of a shard: one per two bytes; size the buffer from a first cheap bound instead
    cap = min(cap, max(1024, nbytes // 8 + 1024))
    fp = torch.empty((cap, 2), dtype=torch.int64, device=device)
    n_rec = op.rmdup_prepare_device(d_in_ptr, nbytes, fp.data_ptr(), cap)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_rec], dtype=torch.int64, device=device), group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    mine = torch.zeros((mx, 2), dtype=torch.int64, device=device)
    mine[:n_rec] = fp[:n_rec]
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    before = [gathered[r][:counts[r]] for r in range(rank)]
    n_before = sum(counts[:rank])
    all_before = torch.cat(before) if n_before else torch.zeros((1, 2), dtype=torch.int64, device=device)
    all_before = all_before.contiguous()
    out = op.rmdup_resolve_device(all_before.data_ptr(), n_before)
    return out, n_rec


def run_sharded(op_name, opts, shard, device_index=-1, lib=None):
    """Map-only operator on this rank's shard (host bytes); returns (Result, global offset, global total)."""
    with Operator(op_name, opts, device=device_index, lib=lib) as op:
        res = op.call(shard, partition_id=_world()[0])
    off, total = output_offsets(len(res.data))
    return res, off, total
