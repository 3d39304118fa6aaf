"""Multi-rank host logic on CPU: world_size 2, gloo backend, kernels through the development emulator build.
Checks that sharded runs + the one exchange step per operator reproduce the single-shard oracle answer."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, data, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import get_lib
        from bigseqkit_b200 import dist as bd
        from bigseqkit_b200.api import Operator
        import ctypes as C
        lib = get_lib("emu")
        cuts = bd.shard_bounds(data, world)
        shard = data[cuts[rank]:cuts[rank + 1]]
        if case == "seq":
            opts = {"Reverse": True, "Complement": True}
            with Operator("SeqTransform", opts, lib=lib) as op:
                res = op.call(shard, partition_id=rank)
            off, total = bd.output_offsets(len(res.data))
            ret[rank] = (off, total, res.data)
        elif case == "stats":
            with Operator("Stats", {"Tabular": True, "All": True}, lib=lib) as op:
                op.call(shard, partition_id=rank)
                bd.stats_allreduce(op)
                ret[rank] = op.stats_render()
        elif case == "rmdup":
            # "device" memory of the emulator is host memory: a CPU tensor stands in for the HBM shard
            t = torch.frombuffer(bytearray(shard) + bytearray(64), dtype=torch.uint8)
            with Operator("RmDup", {"BySeq": True}, lib=lib) as op:
                out, n_rec = bd.rmdup_union(op, t.data_ptr(), len(shard))
                kept = C.string_at(out.data, out.n) if out.n else b""
            off, total = bd.output_offsets(len(kept))
            ret[rank] = (off, total, kept, n_rec)
    finally:
        dist.destroy_process_group()


def _run(case, data, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, data, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


@pytest.fixture(scope="module")
def fastq():
    from bigseqkit_b200 import synth
    return synth.fastq_reads(96 << 10, seed=51, dup_frac=0.3).tobytes()


def test_shard_bounds_are_record_aligned(fastq):
    from bigseqkit_b200 import dist as bd
    import oracle
    starts = set(oracle.frame(fastq))
    for world in (1, 2, 3, 8):
        cuts = bd.shard_bounds(fastq, world)
        assert cuts[0] == 0 and cuts[-1] == len(fastq) and len(cuts) == world + 1
        assert all(c in starts for c in cuts)
    # a quality line that starts with '@' right at the nominal cut is not a record start
    tricky = b"@a\nACGT\n+\n@III\n" * 50
    for world in (2, 3, 5):
        for c in bd.shard_bounds(tricky, world):
            assert c % 15 == 0


def test_seq_two_ranks_ordered_merge(fastq):
    import oracle
    exp, _ = oracle.seq(fastq, {"Reverse": True, "Complement": True})
    parts = _run("seq", fastq)
    merged = bytearray(parts[0][1])
    for off, total, d in parts:
        assert total == len(exp)
        merged[off:off + len(d)] = d
    assert bytes(merged) == exp


def test_stats_two_ranks_allreduce(fastq):
    import oracle
    exp = oracle.stats(fastq, {"Tabular": True, "All": True})[1]
    rows = _run("stats", fastq)
    assert rows[0] == exp and rows[1] == exp


def test_rmdup_two_ranks_fingerprint_union(fastq):
    import oracle
    exp, _, _ = oracle.rmdup(fastq, {"BySeq": True})
    parts = _run("rmdup", fastq)
    merged = bytearray(parts[0][1])
    for off, total, d, _ in parts:
        assert total == len(exp)
        merged[off:off + len(d)] = d
    assert bytes(merged) == exp
    assert sum(p[3] for p in parts) == len(oracle.frame(fastq)) - 1
