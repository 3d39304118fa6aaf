"""seq / subseq / stats: CUDA path (through the C ABI) vs the CPU oracle, bit-exact."""
import pytest

import oracle
from bigseqkit_b200.api import BskError, Operator
from cases import EDGE_INPUTS, FA_SIMPLE, FQ_SIMPLE, fuzz_inputs
from util import check_parity

SEQ_OPTS = [
    {},
    {"Reverse": True, "Complement": True},
    {"Reverse": True},
    {"Complement": True},
    {"Config": {"LineWidth": 4}},
    {"Config": {"LineWidth": 0}},
    {"RemoveGaps": True},
    {"RemoveGaps": True, "GapLetters": "-N"},
    {"Name": True},
    {"Name": True, "OnlyId": True},
    {"Seq": True},
    {"Seq": True, "Reverse": True, "Complement": True},
    {"Qual": True},
    {"Qual": True, "Reverse": True},
    {"Name": True, "Seq": True},
    {"OnlyId": True},
    {"LowerCase": True},
    {"UpperCase": True, "Complement": True},
    {"Dna2rna": True},
    {"Rna2dna": True},
    {"MinLen": 3},
    {"MaxLen": 5},
    {"MinLen": 2, "MaxLen": 100, "Reverse": True},
    {"MinQual": 30.0},
    {"MaxQual": 30.0},
    {"ValidateSeq": True},
    {"Config": {"SeqType": "dna"}},
    {"Config": {"SeqType": "protein"}, "Complement": True},
    {"Config": {"IDNCBI": True}, "OnlyId": True},
    {"Config": {"AlphabetGuessSeqLength": 0}, "Complement": True},
]

BAD_SEQ_OPTS = [
    {"GapLetters": ""},
    {"MinLen": 10, "MaxLen": 5},
    {"MinQual": 30.0, "MaxQual": 10.0},
    {"LowerCase": True, "UpperCase": True},
    {"Config": {"SeqType": "banana"}},
]


@pytest.mark.parametrize("opts", SEQ_OPTS, ids=lambda o: str(o)[:60])
def test_seq_edge_inputs(lib, opts):
    for name, data in EDGE_INPUTS.items():
        check_parity(lib, "SeqTransform", data, opts)


@pytest.mark.parametrize("opts", BAD_SEQ_OPTS, ids=lambda o: str(o)[:60])
def test_seq_bad_flags(lib, opts):
    check_parity(lib, "SeqTransform", FQ_SIMPLE, opts)


@pytest.mark.parametrize("seed", range(6))
def test_seq_fuzz(lib, seed):
    for data in fuzz_inputs(seed):
        for opts in ({}, {"Reverse": True, "Complement": True}, {"Config": {"LineWidth": 11}, "RemoveGaps": True},
                     {"MinLen": 5, "OnlyId": True}):
            check_parity(lib, "SeqTransform", data, opts)


def test_seq_kat(lib):
    # SURVEY 4.3 hand-derived vectors (bigseqkit-lib/seq.go:188-196, helper.go:240-250)
    with Operator("SeqTransform", {"Reverse": True, "Complement": True}, lib=lib) as o:
        assert o.call(b"@r1 d\nACGTN\n+\nIIIJK\n").data == b"@r1 d\nNACGT\n+\nKJIII\n"
    with Operator("SeqTransform", {"Config": {"LineWidth": 4}}, lib=lib) as o:
        assert o.call(b">s1\nACGTAC\nGT\n").data == b">s1\nACGT\nACGT\n"
    with Operator("SeqTransform", {"RemoveGaps": True}, lib=lib) as o:
        assert o.call(b">s1\nAC-G T\n").data == b">s1\nACGT\n"


SUBSEQ_REGIONS = ["1:1", "2:4", "-4:-2", "-4:-1", "-1:-1", "2:-2", "1:-1", "1:12", "-12:-1", "5:3", "100:200", "3:3"]


@pytest.mark.parametrize("region", SUBSEQ_REGIONS)
def test_subseq_regions(lib, region):
    for name, data in EDGE_INPUTS.items():
        check_parity(lib, "SubseqTransform", data, {"Region": region})


def test_subseq_kat(lib):
    # region table of bigseqkit-cli/helper.go:348-361
    table = {"1:1": b"A", "2:4": b"CGT", "-4:-2": b"cgt", "-4:-1": b"cgtn", "-1:-1": b"n", "2:-2": b"CGTNacgt",
             "1:-1": b"ACGTNacgtn", "1:12": b"ACGTNacgtn", "-12:-1": b"ACGTNacgtn"}
    for region, exp in table.items():
        with Operator("SubseqTransform", {"Region": region}, lib=lib) as o:
            assert o.call(b">x\nACGTNacgtn\n").data == b">x\n" + exp + b"\n"


@pytest.mark.parametrize("region", ["0:5", "1:0", "-3:4", "abc", "1-5", ""])
def test_subseq_bad_region(lib, region):
    check_parity(lib, "SubseqTransform", FA_SIMPLE, {"Region": region})


def test_subseq_fuzz(lib):
    for seed in range(3):
        for data in fuzz_inputs(100 + seed):
            for region in ("3:20", "-15:-2", "10:-10"):
                check_parity(lib, "SubseqTransform", data, {"Region": region})


# ------------------------------------------------------------------ stats
STATS_OPTS = [{"Tabular": True}, {"Tabular": True, "All": True}, {"All": True}, {},
              {"Tabular": True, "All": True, "FqEncoding": "illumina-1.3+"},
              {"Tabular": True, "All": True, "GapLetters": "-N"}, {"Tabular": True, "Config": {"SeqType": "protein"}}]


def stats_both(lib, data, opts):
    try:
        exp = oracle.stats(data, opts)
        exp_err = None
    except oracle.OracleError as e:
        exp, exp_err = None, str(e)
    got_err = None
    got = None
    try:
        with Operator("Stats", opts, lib=lib) as o:
            o.call(data)
            got = (o.stats_result(), o.stats_render())
    except BskError as e:
        got_err = str(e)
    assert got_err == exp_err
    if exp is not None:
        assert got[1] == exp[1]
        for k, v in exp[0].items():
            if isinstance(v, float) and v != v:
                assert got[0][k] != got[0][k]
            else:
                assert got[0][k] == v, (k, got[0][k], v)


@pytest.mark.parametrize("opts", STATS_OPTS, ids=lambda o: str(o)[:60])
def test_stats_edge_inputs(lib, opts):
    for name, data in EDGE_INPUTS.items():
        stats_both(lib, data, opts)


def test_stats_fuzz(lib):
    for seed in range(4):
        for data in fuzz_inputs(200 + seed):
            stats_both(lib, data, {"Tabular": True, "All": True})


def test_stats_bad_flags(lib):
    stats_both(lib, FQ_SIMPLE, {"FqEncoding": "bogus"})
    stats_both(lib, FQ_SIMPLE, {"GapLetters": ""})


def test_stats_kat(lib):
    # 3 FASTA records of length 4, 6, 6 with -T (bigseqkit/stats.go:199-207)
    with Operator("Stats", {"Tabular": True}, lib=lib) as o:
        o.call(b">a\nACGT\n>b\nACGTAC\n>c\nACGTAC\n")
        assert o.stats_render() == ("file\tformat\ttype\tnum_seqs\tsum_len\tmin_len\tavg_len\tmax_len\n"
                                    "input0\tN/A\tDNA\t3\t16\t4\t5.3\t6\n")


def test_stats_merge_sum_semantics(lib):
    # StatsReduce with sum semantics: shards merged == whole
    data = fuzz_inputs(7)[1]
    recs = oracle.frame(data)
    cut = recs[len(recs) // 2]
    opts = {"Tabular": True, "All": True}
    exp = oracle.stats_sharded([data[:cut], data[cut:]], opts)
    with Operator("Stats", opts, lib=lib) as a, Operator("Stats", opts, lib=lib) as b:
        a.call(data[:cut])
        b.call(data[cut:])
        a.stats_merge(b)
        assert a.stats_render() == exp[1]
    with Operator("Stats", opts, lib=lib) as a:  # accumulation over successive calls
        a.call(data[:cut])
        a.call(data[cut:])
        assert a.stats_render() == exp[1]
