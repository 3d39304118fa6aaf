"""fq2fa: CUDA path vs the CPU oracle (Fq2Fa.Call, bigseqkit-lib/fq2fa.go:36-61)."""
import random

import pytest

import oracle
from bigseqkit_b200.api import Operator
from cases import EDGE_INPUTS, FA_SIMPLE, FQ_SIMPLE, fuzz_fasta, fuzz_fastq
from util import check_parity


def test_fq2fa_kat(lib):
    # hand-derived: qualities dropped, ">" marker, sequence on one line, FileStore's '\n' after every element
    d = b"@r1 first read\nACGT\n+\nIIII\n@r2\nGGCCA\n+r2\n!!!!!\n"
    with Operator("Fq2Fa", {}, lib=lib) as o:
        r = o.call(d)
    assert r.data == b">r1 first read\nACGT\n>r2\nGGCCA\n"
    assert list(r.elem_off) == [0, 20, 30]
    # FASTA input: lines joined (Format(0)), nothing else changes
    with Operator("Fq2Fa", {}, lib=lib) as o:
        assert o.call(b">a\nAC\nGT\n>b desc\n\nTT\n").data == b">a\nACGT\n>b desc\nTT\n"


def test_fq2fa_edge_inputs(lib):
    for name, data in EDGE_INPUTS.items():
        check_parity(lib, "Fq2Fa", data, {})
    check_parity(lib, "Fq2Fa", FQ_SIMPLE, {})
    check_parity(lib, "Fq2Fa", FA_SIMPLE, {"Config": {"LineWidth": 3}})  # the width is ignored (Format(0))


@pytest.mark.parametrize("seed", range(4))
def test_fq2fa_fuzz(lib, seed):
    rng = random.Random(300 + seed)
    check_parity(lib, "Fq2Fa", fuzz_fastq(rng, n_rec=80, max_len=120), {})
    check_parity(lib, "Fq2Fa", fuzz_fasta(rng, n_rec=60, max_len=200), {})


def test_fq2fa_reads_and_blocks(lib, monkeypatch):
    from bigseqkit_b200 import synth
    data = synth.fastq_reads(200 << 10, seed=91).tobytes()
    exp = oracle.fq2fa(data, {})
    got = check_parity(lib, "Fq2Fa", data, {})
    assert got[0].count(b">") >= len(exp[1]) - 1 and b"\n+\n" not in got[0]
    monkeypatch.setenv("BSK_BLOCK_BYTES", "8192")  # several record-aligned blocks per partition
    check_parity(lib, "Fq2Fa", data, {})
    # a FASTQ length mismatch is the parser's error on both sides
    check_parity(lib, "Fq2Fa", b"@a\nACGT\n+\nII\n", {})
