"""The one-pass stats tile kernel (k_stats_tile.cu): parity with the oracle on short-record inputs, clean hand-over
to the general path for everything else."""
import random

import pytest

import oracle
from bigseqkit_b200 import synth
from bigseqkit_b200.api import Operator
from cases import EDGE_INPUTS, fuzz_fasta, fuzz_fastq
from test_fused_path import _fixed_fastq

OPTS = [{"Tabular": True}, {"Tabular": True, "All": True}, {"All": True, "FqEncoding": "illumina-1.3+"},
        {"Tabular": True, "All": True, "GapLetters": "-N"}]


def inputs():
    rng = random.Random(77)
    return {
        "fastq_reads": synth.fastq_reads(300 << 10, seed=61).tobytes(),
        "fastq_plus_with_name": _fixed_fastq(600, 9, 80, 62, plus="x y"),
        "fastq_varlen": fuzz_fastq(rng, n_rec=1500, max_len=300),
        "fasta_reads": synth.fasta_reads(3000, read_len=100, seed=63).tobytes(),
        "fasta_wrapped": synth.fasta_reads(2000, read_len=150, seed=64, width=60).tobytes(),
        "fasta_cds": synth.fasta_cds(300 << 10, seed=65).tobytes(),
        "fasta_gaps": fuzz_fasta(rng, n_rec=900, max_len=500, alphabet="ACGT-. N"),
        "fasta_no_final_newline": synth.fasta_reads(500, read_len=100, seed=66, width=60).tobytes()[:-1],
        "fastq_no_final_newline": synth.fastq_reads(50 << 10, seed=67).tobytes()[:-1],
        "fastq_short_then_long": _fixed_fastq(1, 4, 20, 68) + _fixed_fastq(120, 10, 900, 69) + _fixed_fastq(40, 10, 1700, 70),
        "fastq_rec48_many_lines": _fixed_fastq(3000, 6, 18, 71),
        "fasta_header_only_records": b">a\n>b\nACGT\n>c\n" * 400,
        "fasta_blank_lines": b">a\nAC\n\nGT\n\n>b\n\nGG\n" * 300,
    }


def run(lib, data, opts):
    with Operator("Stats", opts, lib=lib) as o:
        r = o.call(data)
        return r.n_records, o.stats_render(), o.stats_result(), o.timings()


# fuzzed FASTQ: multi-line / malformed records, general path (or the reference's error); the other two hold more than 2048
# lines per 24 KiB region (12-byte lines), which the 4-CTAs-per-SM shape of the kernel leaves to the general path
MAY_FALL_BACK = {"fastq_varlen", "fastq_rec48_many_lines", "fasta_blank_lines"}


@pytest.mark.parametrize("opts", OPTS, ids=lambda o: str(o)[:50])
def test_stats_tile_parity(lib, opts):
    for name, data in inputs().items():
        try:
            (exp_res, exp_row), exp_err = oracle.stats(data, opts), None
        except oracle.OracleError as e:
            exp_err = str(e)
        try:
            (nrec, row, res, t), err = run(lib, data, opts), None
        except Exception as e:  # noqa: BLE001
            err = str(e)
        assert err == exp_err, (name, opts)
        if exp_err is not None:
            continue
        assert row == exp_row, (name, opts)
        assert res["hist"] == exp_res["hist"], (name, opts)
        assert nrec == len(oracle.frame(data)) - 1, name
        if name not in MAY_FALL_BACK:
            assert t["fused_blocks"] == 1 and t["kernel_launches"] == 1, (name, t)


def test_stats_tile_edge_inputs_any_path(lib):
    for name, data in EDGE_INPUTS.items():
        for opts in OPTS[:2]:
            try:
                exp, exp_err = oracle.stats(data, opts)[1], None
            except oracle.OracleError as e:
                exp, exp_err = None, str(e)
            try:
                got, err = run(lib, data, opts)[1], None
            except Exception as e:  # noqa: BLE001
                got, err = None, str(e)
            assert err == exp_err, (name, opts)
            assert got == exp, (name, opts)


def test_stats_tile_declines_long_records(lib):
    data = b">long\n" + b"ACGTTGCA" * 4000 + b"\n" + synth.fasta_reads(300, read_len=100, seed=72).tobytes()
    nrec, row, res, t = run(lib, data, {"Tabular": True, "All": True})
    assert row == oracle.stats(data, {"Tabular": True, "All": True})[1]
    assert t["fused_blocks"] == 0


def test_stats_tile_multi_block_sum_semantics(lib, monkeypatch):
    data = synth.fastq_reads(200 << 10, seed=73).tobytes()
    monkeypatch.setenv("BSK_BLOCK_BYTES", str(48 << 10))
    nrec, row, res, t = run(lib, data, {"Tabular": True, "All": True})
    assert row == oracle.stats(data, {"Tabular": True, "All": True})[1]
    assert t["fused_blocks"] >= 4
