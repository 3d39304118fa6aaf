"""locate / grep (exact-match paths): CUDA path vs the CPU oracle.

Row order is pinned: record order, pattern order as given, '+' rows ascending, then '-' rows
ascending on the reverse strand; the oracle emits the same order, so the comparison is byte for byte.
"""
import random

import pytest

from bigseqkit_b200.api import Operator
from cases import EDGE_INPUTS, fuzz_fasta, fuzz_fastq
from util import check_parity

LOCATE_OPTS = [
    {"Pattern": ["AC"]},
    {"Pattern": ["AC", "GT", "ACGT", "TTT", "N"]},
    {"Pattern": ["AC"], "OnlyPositiveStrand": True},
    {"Pattern": ["ac", "GG"], "IgnoreCase": True},
    {"Pattern": ["AA", "ACG"], "NonGreedy": True},
    {"Pattern": ["ACGT", "CA"], "Bed": True},
    {"Pattern": ["ACGT", "CA"], "Gtf": True},
    {"Pattern": ["ACGT", "CA"], "HideMatched": True},
    {"Pattern": ["GTAC", "ACGTAC"], "Circular": True},
    {"Pattern": ["AC", "ACGTACGTACGTACGTACGTAC"], "Circular": True, "NonGreedy": True},
    {"Pattern": [["named", "ACG"], ["other", "TTGCA"]]},
    {"Pattern": ["MKV", "LL"], "Config": {"SeqType": "protein"}},
    {"Pattern": ["ACGU", "UU"]},
    {"Pattern": ["AC"], "Config": {"IDNCBI": True}},
]
BAD_LOCATE_OPTS = [{}, {"Pattern": [""]}, {"Pattern": ["AC.T"]}, {"Pattern": ["AC", "A#T"]}, {"Pattern": ["AC", ""]}]


@pytest.mark.parametrize("opts", LOCATE_OPTS, ids=lambda o: str(o)[:70])
def test_locate_edge_inputs(lib, opts):
    for name, data in EDGE_INPUTS.items():
        check_parity(lib, "Locate", data, opts)


@pytest.mark.parametrize("opts", BAD_LOCATE_OPTS, ids=lambda o: str(o)[:70])
def test_locate_bad_flags(lib, opts):
    check_parity(lib, "Locate", b">c\nGATTACA\n", opts)


def test_locate_kat(lib):
    # SURVEY 4.3 (bigseqkit-lib/locate.go:583-766)
    with Operator("Locate", {"Pattern": ["TAC", "GTA"]}, lib=lib) as o:
        r = o.call(b">c\nGATTACA\n")
    assert r.data == (b"seqID\tpatternName\tpattern\tstrand\tstart\tend\tmatched\n"
                      b"c\tTAC\tTAC\t+\t4\t6\tTAC\nc\tGTA\tGTA\t-\t4\t6\tGTA\n")
    with Operator("Locate", {"Pattern": ["AA"], "OnlyPositiveStrand": True}, lib=lib) as o:
        rows = o.call(b">c\nAAAA\n").elements()[1:]
    assert [x.split(b"\t")[4:6] for x in rows] == [[b"1", b"2"], [b"2", b"3"], [b"3", b"4"]]
    with Operator("Locate", {"Pattern": ["AA"], "OnlyPositiveStrand": True, "NonGreedy": True}, lib=lib) as o:
        rows = o.call(b">c\nAAAA\n").elements()[1:]
    assert [x.split(b"\t")[4:6] for x in rows] == [[b"1", b"2"]]


def test_locate_header_only_in_partition_zero(lib):
    # Locate.Call(pid, ...): header row only when pid == 0 (bigseqkit-lib/locate.go:198-204)
    with Operator("Locate", {"Pattern": ["TAC"]}, lib=lib) as o:
        a = o.call(b">c\nGATTACA\n", partition_id=0).data
        b = o.call(b">c\nGATTACA\n", partition_id=3).data
    assert a.startswith(b"seqID\t") and b == b"c\tTAC\tTAC\t+\t4\t6\tTAC\n"


@pytest.mark.parametrize("seed", range(4))
def test_locate_fuzz(lib, seed):
    rng = random.Random(400 + seed)
    pats = ["".join(rng.choice("ACGT") for _ in range(rng.choice([1, 2, 3, 4, 6, 9]))) for _ in range(12)]
    pats = list(dict.fromkeys(pats))
    for data in (fuzz_fasta(rng, n_rec=20, max_len=400), fuzz_fasta(rng, n_rec=3, max_len=9000, width=60),
                 fuzz_fastq(rng, n_rec=20), fuzz_fasta(rng, n_rec=10, alphabet="ACGTacgtN", max_len=300)):
        for extra in ({}, {"IgnoreCase": True}, {"NonGreedy": True}, {"Circular": True}, {"OnlyPositiveStrand": True, "Bed": True}):
            opts = {"Pattern": [p.lower() for p in pats] if extra.get("IgnoreCase") else pats}
            opts.update(extra)
            check_parity(lib, "Locate", data, opts)


def test_locate_long_records_and_many_patterns(lib):
    # contigs longer than one 4096-position tile, patterns of several lengths incl. a palindrome
    rng = random.Random(77)
    contig = "".join(rng.choice("ACGT") for _ in range(20000))
    data = (">big one\n" + "\n".join(contig[i:i + 60] for i in range(0, len(contig), 60)) + "\n>small\nACGTACGT\n").encode()
    pats = [contig[100:112], contig[4090:4102], contig[8185:8200], "ACGT", "GAATTC", contig[19990:20000]]
    pats += ["".join(rng.choice("ACGT") for _ in range(12)) for _ in range(200)]
    pats = list(dict.fromkeys(pats))
    check_parity(lib, "Locate", data, {"Pattern": pats})
    check_parity(lib, "Locate", data, {"Pattern": pats, "Circular": True})


GREP_OPTS = [
    {"Pattern": ["s1", "r2", "b"]},
    {"Pattern": ["s1 x", "a desc"], "ByName": True},
    {"Pattern": ["S1", "R2"], "IgnoreCase": True},
    {"Pattern": ["s1"], "InvertMatch": True},
    {"Pattern": ["s1", "a"], "Count": True},
    {"Pattern": ["ACGT"], "BySeq": True},
    {"Pattern": ["GTA", "TTTT"], "BySeq": True},
    {"Pattern": ["GTA"], "BySeq": True, "OnlyPositiveStrand": True},
    {"Pattern": ["acg"], "BySeq": True, "IgnoreCase": True},
    {"Pattern": ["ACG"], "BySeq": True, "InvertMatch": True, "Count": True},
    {"Pattern": ["CG"], "Region": "1:4"},
    {"Pattern": ["CG"], "Region": "-4:-1"},
    {"Pattern": ["GTAC"], "BySeq": True, "Circular": True},
    {"Pattern": ["MKV"], "BySeq": True, "Config": {"SeqType": "protein"}},
    {"Pattern": ["NC_002516.2"], "Config": {"IDNCBI": True}},
    {"Pattern": ["AC", ""], "BySeq": True},
]
BAD_GREP_OPTS = [{}, {"Pattern": ["A#"], "BySeq": True}, {"Pattern": ["AC"], "Region": "0:3"}, {"Pattern": ["AC"], "Region": "x"}]


@pytest.mark.parametrize("opts", GREP_OPTS, ids=lambda o: str(o)[:70])
def test_grep_edge_inputs(lib, opts):
    for name, data in EDGE_INPUTS.items():
        check_parity(lib, "Grep", data, opts)


@pytest.mark.parametrize("opts", BAD_GREP_OPTS, ids=lambda o: str(o)[:70])
def test_grep_bad_flags(lib, opts):
    check_parity(lib, "Grep", b">x\nGATTACA\n", opts)


def test_grep_kat(lib):
    # exact whole-ID set lookup (grep.go:501-512) and '-' strand hit (grep.go:442-482)
    with Operator("Grep", {"Pattern": ["b"]}, lib=lib) as o:
        assert o.call(b">a\nAC\n>b\nGG\n>c\nTT\n").data == b">b\nGG\n"
    with Operator("Grep", {"Pattern": ["GTA"], "BySeq": True}, lib=lib) as o:
        assert o.call(b">x\nGATTACA\n").data == b">x\nGATTACA\n"
        assert o.grep_count() == 1


@pytest.mark.parametrize("seed", range(3))
def test_grep_fuzz(lib, seed):
    rng = random.Random(500 + seed)
    for data in (fuzz_fasta(rng, n_rec=30, max_len=200), fuzz_fastq(rng, n_rec=30), fuzz_fasta(rng, n_rec=4, max_len=6000, width=60)):
        ids = ["s%d" % rng.randint(0, 30) for _ in range(6)] + ["r%d" % rng.randint(0, 30) for _ in range(6)]
        pats = ["".join(rng.choice("ACGT") for _ in range(rng.choice([2, 3, 5]))) for _ in range(4)]
        for opts in ({"Pattern": ids}, {"Pattern": ids, "InvertMatch": True}, {"Pattern": pats, "BySeq": True},
                     {"Pattern": pats, "BySeq": True, "InvertMatch": True}, {"Pattern": pats, "Region": "5:40"},
                     {"Pattern": pats, "Region": "-30:-3", "Count": True}):
            check_parity(lib, "Grep", data, opts)
