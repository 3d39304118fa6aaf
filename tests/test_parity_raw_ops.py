"""duplicate / range / head: operators on the raw elements (bigseqkit-lib/duplicate.go:24-30, range.go:26-44,
bigseqkit/head.go:33-44) against the oracle, on the shared edge inputs, fuzz inputs and many-block partitions."""
import pytest

import oracle
from bigseqkit_b200 import synth
from bigseqkit_b200.api import Operator
from cases import EDGE_INPUTS, fuzz_inputs

# inputs the framing rule accepts as they stand (a first byte that is no marker is an error of the parsing operators,
# not of these raw ones: the general index reports it; they are covered by the seq tests)
RAW_INPUTS = {k: v for k, v in EDGE_INPUTS.items() if k not in ("fa_no_marker_first", "fa_leading_newline")}


def _run(lib, name, opts, data):
    with Operator(name, opts, lib=lib) as op:
        r = op.call(data)
        return r.data, list(r.elem_off)


@pytest.mark.parametrize("times", [0, 1, 2, 5])
def test_duplicate(lib, times):
    for name, data in list(RAW_INPUTS.items()) + [("fuzz%d" % i, d) for i, d in enumerate(fuzz_inputs(31, 8))]:
        exp = oracle.duplicate(data, times)
        got = _run(lib, "Duplicate", {"Times": times}, data)
        assert got[0] == exp[0], (name, times)
        if times:
            assert got[1] == exp[1], (name, times)


@pytest.mark.parametrize("rng", [(0, 1), (0, 3), (1, 2), (2, 1 << 40), (5, 5), (3, 2), (0, 0), (100000, 100010)], ids=str)
def test_range(lib, rng):
    for name, data in list(RAW_INPUTS.items()) + [("fuzz%d" % i, d) for i, d in enumerate(fuzz_inputs(32, 8))]:
        exp = oracle.range_(data, rng[0], rng[1])
        got = _run(lib, "RangePrepare", {"start": rng[0], "end": rng[1]}, data)
        assert got[0] == exp[0], (name, rng)
        if exp[0]:
            assert got[1] == exp[1], (name, rng)


def test_head_and_index_base(lib):
    data = synth.fastq_reads(64 << 10, seed=5).tobytes()
    for n in (1, 10, 10**6):
        assert _run(lib, "Head", {"N": n}, data)[0] == oracle.head(data, n)[0]
    # a later partition of the same dataframe: MapWithIndex keeps counting
    assert _run(lib, "Range", {"Start": 1000, "End": 1003, "IndexBase": 990}, data)[0] == oracle.range_(data, 1000, 1003, 990)[0]


def test_many_blocks(lib, monkeypatch):
    # the record index keeps running from block to block of one partition
    monkeypatch.setenv("BSK_BLOCK_BYTES", "8192")
    fq = synth.fastq_reads(100 << 10, seed=6).tobytes()[:-1]  # no final newline
    fa = synth.fasta_cds(80 << 10, seed=7).tobytes()
    for data in (fq, fa):
        for a, b in ((0, 7), (40, 90), (250, 1 << 50)):
            exp = oracle.range_(data, a, b)
            got = _run(lib, "Range", {"Start": a, "End": b}, data)
            assert got[0] == exp[0] and got[1] == exp[1], (a, b)
        exp = oracle.duplicate(data, 3)
        got = _run(lib, "Duplicate", {"Times": 3}, data)
        assert got[0] == exp[0] and got[1] == exp[1]
