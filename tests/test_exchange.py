"""The two exchange steps behind the C ABI, single-process form (bsk_reduce / bsk_rmdup_union): several ctxs of one
process, ctx order == input order.  Runs on the emulator here and on ONE GPU on the box (all ctxs on device 0), so the
exchange logic is exercised even where a second GPU for the NCCL form (tests/test_dist_nccl.py) is missing.

  StatsReduce  bigseqkit-lib/stats.go:128-137, folded by Reduce (bigseqkit/stats.go:91)   sum semantics
  GroupByKey + RmDupCheck  bigseqkit/rmdup.go:97, bigseqkit-lib/rmdup.go:118-242          first in input order wins (Q4)
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from bigseqkit_b200 import dist as bd
from bigseqkit_b200 import synth
from bigseqkit_b200.api import Operator, reduce_local, rmdup_union_local


def _to_device(lib, data):
    """bytes -> (keepalive, device pointer).  The emulator's device memory is host memory."""
    if "emu" in lib.path:
        buf = np.frombuffer(bytearray(data) + bytearray(64), dtype=np.uint8)
        return buf, buf.ctypes.data
    import torch
    t = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()
    return t, t.data_ptr()


@pytest.mark.parametrize("world", [2, 3, 5])
def test_stats_reduce_local(lib, world):
    data = synth.fastq_reads(160 << 10, seed=61).tobytes()
    cuts = bd.shard_bounds(data, world)
    sopts = {"Tabular": True, "All": True}
    ops = [Operator("Stats", sopts, lib=lib) for _ in range(world)]
    try:
        for r, op in enumerate(ops):
            op.call(data[cuts[r]:cuts[r + 1]], partition_id=r)
        reduce_local(ops)
        exp = oracle.stats(data, sopts)[1]
        for op in ops:
            assert op.stats_render() == exp
    finally:
        for op in ops:
            op.close()


@pytest.mark.parametrize("world,opts", [(2, {"BySeq": True}), (3, {"BySeq": True}), (4, {"ByName": True}), (2, {})])
def test_rmdup_union_local(lib, world, opts):
    data = synth.fastq_reads(200 << 10, seed=62, dup_frac=0.3).tobytes()
    if not opts.get("BySeq"):  # duplicate some names / ids as well: repeat a slice of whole records
        st = oracle.frame(data)
        data = data + data[st[3]:st[40]]
    cuts = bd.shard_bounds(data, world)
    ops = [Operator("RmDup", opts, lib=lib) for _ in range(world)]
    keep = []
    try:
        ptrs, sizes = [], []
        for r in range(world):
            k, p = _to_device(lib, data[cuts[r]:cuts[r + 1]])
            keep.append(k)
            ptrs.append(p)
            sizes.append(cuts[r + 1] - cuts[r])
        outs = rmdup_union_local(ops, ptrs, sizes)
        got = b""
        for op, out in zip(ops, outs):
            d, _ = op.fetch(out)
            got += d.tobytes()
        exp, _, removed = oracle.rmdup(data, opts)
        assert removed > 0
        assert got == exp
    finally:
        for op in ops:
            op.close()


def test_rmdup_union_empty_and_single(lib):
    data = synth.fastq_reads(8 << 10, seed=63, dup_frac=0.5).tobytes()
    ops = [Operator("RmDup", {"BySeq": True}, lib=lib) for _ in range(3)]
    try:
        k0, p0 = _to_device(lib, data)
        k1, p1 = _to_device(lib, b"")
        k2, p2 = _to_device(lib, data)
        outs = rmdup_union_local(ops, [p0, p1, p2], [len(data), 0, len(data)])
        exp, _, _ = oracle.rmdup(data, {"BySeq": True})
        assert ops[0].fetch(outs[0])[0].tobytes() == exp
        assert outs[1].n == 0 and outs[2].n == 0  # the third shard repeats the first: everything is a duplicate
    finally:
        for op in ops:
            op.close()


def test_comm_calls_without_communicator(lib):
    with Operator("Stats", {}, lib=lib) as op:
        assert op.comm_rank() == (0, 1)
        assert op.output_offsets(123) == (0, 123)
        with pytest.raises(Exception):
            op.stats_allreduce()
