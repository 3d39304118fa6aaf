"""The C oracle against the independent pure-Python restatement (tests/pytwin.py) on small inputs (CPU only)."""
import random

import oracle
import pytwin
from bigseqkit_b200 import synth
from cases import EDGE_INPUTS, fuzz_fasta

SIMPLE = ["fq_simple", "fa_simple", "fq_no_final_newline", "fq_qual_starts_with_at", "fq_plus_with_name", "fq_multiline",
          "fq_one_base", "fq_lower_iupac", "fa_single", "fa_wrapped", "fa_long_line"]


def inputs():
    rng = random.Random(5)
    d = {k: EDGE_INPUTS[k] for k in SIMPLE}
    d["synth_fastq"] = synth.fastq_reads(40 << 10, seed=3, dup_frac=0.3).tobytes()
    d["synth_cds"] = synth.fasta_cds(40 << 10, seed=4).tobytes()
    d["fuzz_fasta"] = b">first\nACGTACGT\n" + fuzz_fasta(rng, n_rec=60, alphabet="ACGTN", max_len=200)
    return d


def test_framing_agrees():
    for name, data in inputs().items():
        assert oracle.frame(data)[:-1] == pytwin.frame(data), name


def test_seq_revcomp_agrees():
    for name, data in inputs().items():
        assert oracle.seq(data, {"Reverse": True, "Complement": True})[0] == pytwin.seq_revcomp(data), name


def test_stats_row_agrees():
    for name, data in inputs().items():
        if name == "fq_lower_iupac":  # second record is all N: still DNA; keep the type column out of this check
            pass
        row = oracle.stats(data, {"Tabular": True})[1].split("\n")[1]
        assert row.split("\t")[3:] == pytwin.stats_row(data).split("\t")[3:], name


def test_rmdup_by_seq_and_keys_agree():
    for name, data in inputs().items():
        exp, keys = pytwin.rmdup_by_seq(data)
        assert oracle.rmdup(data, {"BySeq": True})[0] == exp, name
        assert list(oracle.rmdup_keys(data, {"BySeq": True})) == keys, name


def test_translate_frame1_agrees():
    for name in ("synth_cds", "fuzz_fasta", "fa_wrapped", "fa_long_line"):
        data = inputs()[name]
        try:
            got = oracle.translate(data, {"AllowUnknownCodon": True})[0]
        except oracle.OracleError:
            continue  # sequences shorter than one codon abort the reference
        assert got == pytwin.translate_frame1(data), name
