"""Multi-GPU plumbing: one process per GPU, shards partition one per rank.

The map-only operators (seq / subseq / translate / grep / locate) need no data-path collective: every rank
runs its record-aligned shard and `output_offsets` (one all-gather of sizes) tells it where its bytes go in the
merged output -- the job the reference does with an MPI token ring in FileStore
(bigseqkit-lib/helper.go:399-431).  Two operators have a real exchange step:

  stats   StatsReduce (bigseqkit-lib/stats.go:128-137, bigseqkit/stats.go:91)   -> one all-reduce / all-gather
  rmdup   GroupByKey  (bigseqkit/rmdup.go:97)                                   -> one all-gather of 16-byte
          fingerprints; the first occurrence in GLOBAL input order survives

The exchange steps live behind the C ABI (bsk_comm_init / bsk_stats_allreduce / bsk_rmdup_sharded /
bsk_output_offsets, csrc/comm.cu: NCCL on the ctx stream).  This module is the thin caller: `init_comm` ships the
NCCL unique id through torch.distributed (the job MPI / an IgnisHPC variable does for the Go shim), after which the
functions below are one C call each.  An operator WITHOUT a communicator (the CPU tests: gloo + the host emulator,
which has no NCCL) takes the `torch.distributed` path with the same semantics, tensors only, no pickling.
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


def _sync(device):
    """libbsk launches on its own non-blocking stream: torch work on `device` must have finished before a buffer
    torch produced is handed to it (and libbsk's entry points synchronise their stream before returning)."""
    if str(device).startswith("cuda"):
        torch.cuda.synchronize(device)


def init_comm(op, group=None):
    """Bind a NCCL communicator to the operator's ctx: rank 0 creates the unique id, torch.distributed ships it."""
    from .api import comm_unique_id
    rank, world = _world(group)
    if world == 1:
        return op
    box = [comm_unique_id(op.lib) if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    op.comm_init(box[0], world, rank)
    return op


def _has_comm(op):
    return op.comm_rank()[1] > 1


def output_offsets(n_local, device="cpu", group=None, op=None):
    """Global byte offset of this rank's output and the total size: all-gather of one int64 per rank."""
    rank, world = _world(group)
    if world == 1:
        return 0, int(n_local)
    if op is not None and _has_comm(op):
        return op.output_offsets(int(n_local))
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, mine, group=group) if str(device).startswith("cuda") else \
        dist.all_gather(list(sizes.split(1)), mine, group=group)
    sizes = [int(x) for x in sizes.tolist()]
    return sum(sizes[:rank]), sum(sizes)


def stats_allreduce(op, device="cpu", nbins=65536, group=None):
    """Merge the Stats totals of all ranks into every rank's operator (sum semantics)."""
    rank, world = _world(group)
    if world == 1:
        return op
    if _has_comm(op):
        op.stats_allreduce()
        return op
    # ---- torch.distributed path (no communicator on the ctx): dense histogram through one all-reduce, the scalars
    # and the type column through one all-gather, lengths >= nbins as padded pairs through another
    res = op.stats_result()
    hist = torch.zeros(nbins + 3, dtype=torch.int64, device=device)
    _sync(device)
    op.stats_dense_device(hist.data_ptr(), nbins)
    hist[nbins:] = torch.tensor([res["q20"], res["q30"], res["sum_gap"]], dtype=torch.int64, device=device)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    longs = [(l, c) for l, c in res["hist"] if l >= nbins]
    info = torch.zeros(4, dtype=torch.int64, device=device)
    info[0], info[1] = res["num"], len(longs)
    tb = res["type"].encode()[:15]
    info[2] = int.from_bytes(tb[:8].ljust(8, b"\0"), "little", signed=True)
    info[3] = int.from_bytes(tb[8:].ljust(8, b"\0"), "little", signed=True)
    infos = [torch.zeros_like(info) for _ in range(world)]
    dist.all_gather(infos, info, group=group)
    infos = [t.tolist() for t in infos]
    max_long = max(i[1] for i in infos)
    pairs = []
    if max_long:
        mine = torch.zeros((max_long, 2), dtype=torch.int64, device=device)
        if longs:
            mine[:len(longs)] = torch.tensor(longs, dtype=torch.int64, device=device)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        for r in range(world):
            pairs.extend((int(l), int(c)) for l, c in gathered[r][:infos[r][1]].tolist())
    h = hist.cpu().numpy()
    merged = {int(l): int(h[l]) for l in np.nonzero(h[:nbins])[0]}
    for l, c in pairs:
        merged[l] = merged.get(l, 0) + c
    # the type column comes from the first rank that saw a record (reference: partition 0's first record)
    typ = ""
    for i in infos:
        if i[0] > 0:
            typ = (i[2].to_bytes(8, "little", signed=True) + i[3].to_bytes(8, "little", signed=True)).rstrip(b"\0").decode()
            break
    op.reset()
    op.stats_add(sorted(merged.items()), q20=int(h[nbins]), q30=int(h[nbins + 1]), sum_gap=int(h[nbins + 2]), type=typ)
    return op


def rmdup_union(op, d_in_ptr, nbytes, device="cpu", group=None):
    """rmdup over all ranks' shards (rank order == input order): hash locally, all-gather the 16-byte fingerprints,
    drop every local record whose fingerprint occurs earlier in global order.  Returns the bsk_out of the survivors
    (device pointers) and the number of local records."""
    rank, world = _world(group)
    if world == 1:
        out = op.call_device(d_in_ptr, nbytes)
        return out, int(out.n_records)
    _sync(device)  # the shard may have been produced by torch on another stream
    if _has_comm(op):
        out = op.rmdup_sharded(d_in_ptr, nbytes)
        return out, int(out.n_records)
    # ---- torch.distributed path.  A record takes at least 5 bytes (">a\nA\n"): a true bound on the shard's records,
    # so no rank can fail on the buffer size while the others wait in the collective.
    cap = nbytes // 5 + 16
    fp = torch.empty((cap, 2), dtype=torch.int64, device=device)
    _sync(device)
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
    _sync(device)  # gather + cat ran on torch's stream; the resolve kernels run on the ctx stream
    out = op.rmdup_resolve_device(all_before.data_ptr(), n_before)
    return out, n_rec


def run_sharded(op_name, opts, shard, device_index=-1, lib=None):
    """Map-only operator on this rank's shard (host bytes); returns (Result, global offset, global total)."""
    with Operator(op_name, opts, device=device_index, lib=lib) as op:
        res = op.call(shard, partition_id=_world()[0])
    off, total = output_offsets(len(res.data))
    return res, off, total
