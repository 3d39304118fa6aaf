"""Pins: the oracle AND the CUDA library against tests/golden/pins.json (made by tests/golden/make_golden.py
from the reference's help-text tables, the Python xxhash package and the hand-derived SURVEY 4.3 vectors)."""
import json
import os

import pytest

import oracle
from bigseqkit_b200.api import Operator

PINS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pins.json")))


# ---------------------------------------------------------------- oracle (CPU, runs without a GPU)
def test_oracle_region_table():
    seq = PINS["region"]["seq"]
    for region, exp in PINS["region"]["regions"].items():
        a, b = (int(x) for x in region.split(":"))
        s0, ln = oracle.subseq_range(len(seq), a, b)
        assert seq[s0:s0 + ln] == exp, region
        out, _ = oracle.subseq((">x\n%s\n" % seq).encode(), {"Region": region})
        assert out == (">x\n%s\n" % exp).encode()


def test_oracle_ambiguous_codons_and_table_ids():
    t = PINS["translate"]
    for codon, aa in t["ambiguous_codons"].items():
        assert oracle.translate_codon(t["table"], codon) == aa, codon
    for tid in t["table_ids"]:
        assert oracle.translate_codon(tid, "GGG") == "G"  # glycine in every NCBI code: table exists
    for tid in (0, 7, 8, 15, 17, 32):
        with pytest.raises(oracle.OracleError):
            oracle.translate(b">x\nATG\n", {"TranslTable": tid})
    bases = "TCAG"
    for i, aa in enumerate(PINS["ncbi_table_1"]):
        codon = bases[i // 16] + bases[(i // 4) % 4] + bases[i % 4]
        assert oracle.translate_codon(1, codon) == aa, codon


def test_oracle_xxh64_vectors():
    for v in PINS["xxh64"]["vectors"]:
        assert "%016x" % oracle.xxh64(v["subject"].encode()) == v["u64"], len(v["subject"])
    # keys as the reference casts them: int64(xxhash.Sum64(seq))
    recs = "".join("@r%d\n%s\n+\n%s\n" % (i, v["subject"], "I" * len(v["subject"])) for i, v in enumerate(PINS["xxh64"]["vectors"]))
    keys = oracle.rmdup_keys(recs.encode(), {"BySeq": True})
    assert list(keys) == [v["go_int64"] for v in PINS["xxh64"]["vectors"]]


@pytest.mark.parametrize("kat", PINS["records"], ids=lambda k: "%s-%s" % (k["op"], k["cite"]))
def test_oracle_record_vectors(kat):
    fn = {"SeqTransform": oracle.seq, "Translate": oracle.translate, "Locate": oracle.locate, "Grep": oracle.grep,
          "SubseqTransform": oracle.subseq, "RmDup": lambda d, o: oracle.rmdup(d, o)[:2]}[kat["op"]]
    out, _ = fn(kat["in"].encode(), kat["opts"])
    assert out == kat["out"].encode()


def test_oracle_stats_vectors():
    for kat in PINS["stats"]:
        assert oracle.stats(kat["in"].encode(), kat["opts"])[1] == kat["out"]


# ---------------------------------------------------------------- CUDA library (emulator here, GPU with -m gpu)
@pytest.mark.parametrize("kat", PINS["records"], ids=lambda k: "%s-%s" % (k["op"], k["cite"]))
def test_lib_record_vectors(lib, kat):
    with Operator(kat["op"], kat["opts"], lib=lib) as o:
        assert o.call(kat["in"].encode()).data == kat["out"].encode()


def test_lib_region_table(lib):
    seq = PINS["region"]["seq"]
    for region, exp in PINS["region"]["regions"].items():
        with Operator("SubseqTransform", {"Region": region}, lib=lib) as o:
            assert o.call((">x\n%s\n" % seq).encode()).data == (">x\n%s\n" % exp).encode()


def test_lib_ambiguous_codons(lib):
    for codon, aa in PINS["translate"]["ambiguous_codons"].items():
        with Operator("Translate", {}, lib=lib) as o:
            assert o.call((">x\n%s\n" % codon).encode()).data == (">x\n%s\n" % aa).encode()


def test_lib_xxh64_keys(lib):
    vs = PINS["xxh64"]["vectors"]
    recs = "".join("@r%d\n%s\n+\n%s\n" % (i, v["subject"], "I" * len(v["subject"])) for i, v in enumerate(vs))
    with Operator("RmDupPrepare", {"BySeq": True}, lib=lib) as o:
        o.call(recs.encode())
        assert list(o.rmdup_keys()) == [v["go_int64"] for v in vs]


def test_lib_stats_vectors(lib):
    for kat in PINS["stats"]:
        with Operator("Stats", kat["opts"], lib=lib) as o:
            o.call(kat["in"].encode())
            assert o.stats_render() == kat["out"]
