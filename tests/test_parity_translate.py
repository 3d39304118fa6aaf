"""translate: CUDA path vs the CPU oracle, plus the reference's own ambiguous-codon table."""
import random

import pytest

import oracle
from bigseqkit_b200.api import Operator
from cases import EDGE_INPUTS, fuzz_fasta
from util import check_parity

TR_OPTS = [
    {}, {"Frame": ["6"]}, {"Frame": ["1", "-1"]}, {"Frame": ["2", "3", "-2", "-3"]}, {"Frame": ["6"], "Trim": True},
    {"Frame": ["6"], "Clean": True}, {"Frame": ["6"], "AllowUnknownCodon": True},
    {"Frame": ["6"], "InitCodonAsM": True, "TranslTable": 11}, {"Frame": ["6"], "AppendFrame": True},
    {"Frame": ["1"], "TranslTable": 2, "Config": {"LineWidth": 10}}, {"Frame": ["6"], "Config": {"LineWidth": 0}},
    {"Frame": ["-1"], "AppendFrame": True, "Config": {"IDNCBI": True}}, {"Frame": ["1", "1", "6", "2"]},
    {"Frame": ["6"], "AllowUnknownCodon": True, "Trim": True, "Clean": True, "Config": {"SeqType": "dna"}},
]
BAD_TR_OPTS = [{"TranslTable": 7}, {"Frame": ["4"]}, {"Frame": ["x"]}, {"TranslTable": 0}]


@pytest.mark.parametrize("opts", TR_OPTS, ids=lambda o: str(o)[:70])
def test_translate_edge_inputs(lib, opts):
    for name, data in EDGE_INPUTS.items():
        check_parity(lib, "Translate", data, opts)


@pytest.mark.parametrize("opts", BAD_TR_OPTS, ids=lambda o: str(o)[:70])
def test_translate_bad_flags(lib, opts):
    check_parity(lib, "Translate", b">x\nATGGCCTAA\n", opts)


@pytest.mark.parametrize("table", [1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 16, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31])
def test_translate_all_tables(lib, table):
    rng = random.Random(table)
    data = fuzz_fasta(rng, n_rec=12, alphabet="ACGTacgtNRYKMSWBDHVU", max_len=400, width=60)
    data = b">lead\nATGGCCATTGTAATGGGCCGCTGAAAGGGTGCCCGATAG\n" + data
    check_parity(lib, "Translate", data, {"Frame": ["6"], "TranslTable": table, "AllowUnknownCodon": True,
                                          "InitCodonAsM": True})


def test_translate_fuzz(lib):
    for seed in range(4):
        rng = random.Random(300 + seed)
        for alphabet in ("ACGT", "ACGTN", "ACGU", "ACGT-"):
            data = b">first\nATGACGTTT\n" + fuzz_fasta(rng, n_rec=25, alphabet=alphabet, max_len=300)
            for opts in ({"Frame": ["6"], "AllowUnknownCodon": True}, {"Frame": ["6"]}, {"Frame": ["3"], "Trim": True}):
                check_parity(lib, "Translate", data, opts)


def test_translate_kat(lib):
    # >x ATGGCCTAA -f 6
    with Operator("Translate", {"Frame": ["6"]}, lib=lib) as o:
        r = o.call(b">x\nATGGCCTAA\n")
    assert r.data == b">x\nMA*\n>x\nWP\n>x\nGL\n>x\nLGH\n>x\n*A\n>x\nRP\n"
    # ambiguous-codon examples of the reference help text (bigseqkit-cli/translate.go:42-52), standard table
    table = {"ACN": "T", "CCN": "P", "CGN": "R", "CTN": "L", "GCN": "A", "GGN": "G", "GTN": "V", "TCN": "S", "MGR": "R",
             "YTR": "L"}
    for codon, aa in table.items():
        with Operator("Translate", {}, lib=lib) as o:
            r = o.call((">c\nATG%s\n" % codon).encode())
        assert r.data == (">c\nM%s\n" % aa).encode(), codon
        assert oracle.translate_codon(1, codon) == aa


def test_translate_wrapped_fasta_both_squeeze_paths(lib, monkeypatch):
    # uniformly wrapped records take the arithmetic squeeze (k_squeeze_uniform), ragged ones the per-line copy
    from bigseqkit_b200 import synth
    from util import run_lib, run_oracle
    uniform = synth.fasta_cds(200 << 10, seed=91).tobytes()
    ragged = uniform.replace(b"\n", b"\n\n", 1) + b">r\nACGTACGTAC\nACG\nACGTACGTACGT\n"
    long_last = b">a\nATGGCC\nTAA\n>b\nATG\nGCCTAA\n"  # second record: last line longer than the first -> ragged
    for name, data in (("uniform", uniform), ("uniform_no_final_newline", uniform[:-1]), ("ragged", ragged), ("long_last", long_last)):
        exp = run_oracle("Translate", data, {"Frame": ["6"]})
        for env in (None, "1"):
            if env:
                monkeypatch.setenv("BSK_NO_UNIFORM_SQUEEZE", env)
            got = run_lib(lib, "Translate", data, {"Frame": ["6"]})
            if env:
                monkeypatch.delenv("BSK_NO_UNIFORM_SQUEEZE")
            assert got[0] == exp[0] and list(got[1]) == list(exp[1]), (name, env)


def test_more_than_66_frames_is_refused(lib):
    # the device-side frame list holds 66 entries; a longer list is an explicit error, not a silently shorter output
    from bigseqkit_b200.api import BskError
    with pytest.raises(BskError):
        Operator("Translate", {"Frame": ["1"] * 67}, lib=lib)
    with Operator("Translate", {"Frame": ["1", "2"] * 33}, lib=lib) as op:
        data = b">a\nATGGCCAAATAA\n"
        assert op.call(data).data == oracle.translate(data, {"Frame": ["1", "2"] * 33})[0]
