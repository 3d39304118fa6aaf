"""Parity at config scale (-m gpu): the WHOLE output of every BASELINE workload block against the oracle's output on the
same block (bytes + element offsets, or the stats row), and bsk_run_buffer across its default 64 MiB pipeline blocks.
The oracle runs sharded over the host threads (orc_run_mt_out); inputs come from the native generators (synth_native.c).
"""
import os

import numpy as np
import pytest

import oracle
from bigseqkit_b200 import synth
from bigseqkit_b200.api import Operator

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1
GIB = 1 << 30

CASES = {
    "seq_rc": ("SeqTransform", {"Reverse": True, "Complement": True}, "seq", lambda n: synth.native_fastq(n, seed=2)),
    "stats_fasta": ("Stats", {"Tabular": True}, "stats", lambda n: synth.native_fasta_reads(n, seed=1)),
    "stats_all": ("Stats", {"Tabular": True, "All": True}, "stats", lambda n: synth.native_fastq(n, seed=2)),
    "rmdup": ("RmDup", {"BySeq": True}, "rmdup", lambda n: synth.native_fastq(n, seed=3, dup_frac=0.2)),
    "translate6": ("Translate", {"Frame": ["6"]}, "translate", lambda n: synth.native_cds(n, seed=5)),
    "locate": ("Locate", {"Pattern": synth.pattern_panel(1000, 12, 40)}, "locate", lambda n: synth.native_contigs(min(n, 256 << 20), seed=4)),
    "fq2fa": ("Fq2Fa", {}, "fq2fa", lambda n: synth.native_fastq(n, seed=2)),
    "subseq": ("SubseqTransform", {"Region": "10:-10"}, "subseq", lambda n: synth.native_fastq(n, seed=2)),
}


def _to_device(arr):
    import torch
    t = torch.empty(arr.nbytes + 64, dtype=torch.uint8, device="cuda")
    t[:arr.nbytes] = torch.from_numpy(arr).cuda()
    torch.cuda.synchronize()
    return t


def _compare(op, out, exp, is_stats):
    if is_stats:
        assert op.stats_render() == exp["row"]
        return
    data, offs = op.fetch(out)
    assert data.nbytes == exp["data"].nbytes
    assert np.array_equal(data, exp["data"])
    assert np.array_equal(offs, exp["elem_off"])


@pytest.mark.parametrize("case", list(CASES))
def test_whole_block_device(case):
    """one bsk_run_device call on a 1 GiB block (locate: 256 MiB)"""
    opn, opts, orc, gen = CASES[case]
    arr, n_rec = gen(GIB)
    exp = oracle.run_mt_full(orc, arr.ctypes.data, arr.nbytes, opts, THREADS)
    t = _to_device(arr)
    with Operator(opn, opts, device=0) as op:
        out = op.call_device(t.data_ptr(), arr.nbytes)
        _compare(op, out, exp, opn == "Stats")


@pytest.mark.parametrize("case", ["seq_rc", "stats_all", "rmdup", "translate6", "locate", "fq2fa"])
def test_run_buffer_default_blocks(case):
    """bsk_run_buffer on 200 MiB of host memory: three default 64 MiB pipeline blocks + a tail, record-aligned cuts"""
    opn, opts, orc, gen = CASES[case]
    arr, n_rec = gen(200 << 20)
    exp = oracle.run_mt_full(orc, arr.ctypes.data, arr.nbytes, opts, THREADS)
    assert "BSK_BLOCK_BYTES" not in os.environ
    with Operator(opn, opts, device=0) as op:
        res = op.call((arr.ctypes.data, arr.nbytes))
        if opn == "Stats":
            assert op.stats_render() == exp["row"]
        else:
            assert len(res.data) == exp["data"].nbytes
            assert res.data == exp["data"].tobytes()
            assert np.array_equal(np.asarray(res.elem_off, dtype=np.uint64), exp["elem_off"])
