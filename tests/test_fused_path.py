"""The single-pass tile path (k_fused.cu) for short records: parity with the oracle on inputs that span
many tiles, and clean fall-back to the general path for everything the tile scheme cannot represent."""
import random

import pytest

import oracle
from bigseqkit_b200 import synth
from bigseqkit_b200.api import Operator
from cases import fuzz_fasta, fuzz_fastq

OPTS = [
    {"Reverse": True, "Complement": True},
    {},
    {"Name": True},
    {"Name": True, "OnlyId": True},
    {"Seq": True, "Reverse": True},
    {"Qual": True},
    {"OnlyId": True, "LowerCase": True},
    {"Config": {"LineWidth": 37}, "Complement": True},
    {"Config": {"LineWidth": 0}},
    {"MinLen": 120, "MaxLen": 2000, "Reverse": True, "Complement": True},
    {"MinLen": 100000},
    {"Dna2rna": True, "UpperCase": True},
]


def run(lib, data, opts):
    with Operator("SeqTransform", opts, lib=lib) as o:
        r = o.call(data)
        return r, o.timings()


def big_inputs():
    rng = random.Random(9)
    return {
        "fastq_reads": synth.fastq_reads(300 << 10, seed=11).tobytes(),
        "fastq_varlen": fuzz_fastq(rng, n_rec=1500, max_len=300),
        "fasta_reads": synth.fasta_reads(2500, read_len=100, seed=3).tobytes(),
        "fasta_wrapped_reads": synth.fasta_reads(1500, read_len=150, seed=4, width=60).tobytes(),
        "fasta_cds": synth.fasta_cds(260 << 10, seed=5).tobytes(),
        "fasta_mixed": fuzz_fasta(rng, n_rec=800, max_len=600, alphabet="ACGTacgtN"),
        "fastq_no_final_newline": synth.fastq_reads(40 << 10, seed=12).tobytes()[:-1],
    }


@pytest.mark.parametrize("opts", OPTS, ids=lambda o: str(o)[:60])
def test_fused_parity_many_tiles(lib, opts):
    for name, data in big_inputs().items():
        try:
            exp = oracle.seq(data, opts)
        except oracle.OracleError:
            continue  # e.g. -q on FASTA: covered by the general-path tests
        r, t = run(lib, data, opts)
        assert r.data == exp[0], (name, opts)
        assert list(r.elem_off) == exp[1], (name, opts)
        assert t["fused_blocks"] == 1, (name, opts, "expected the single-pass tile path")


def test_general_path_when_fused_disabled(lib, monkeypatch):
    monkeypatch.setenv("BSK_NO_FUSED", "1")
    data = synth.fastq_reads(100 << 10, seed=13).tobytes()
    opts = {"Reverse": True, "Complement": True}
    r, t = run(lib, data, opts)
    exp = oracle.seq(data, opts)
    assert r.data == exp[0] and list(r.elem_off) == exp[1]
    assert t["fused_blocks"] == 0


FALLBACK_INPUTS = {
    "fq_multiline": b"@a\nACGT\nAC\n+\nIIII\nII\n@b\nGG\n+\nJJ\n",
    "fq_mismatch": b"@a\nACGT\n+\nIII\n@b\nGG\n+\nJJ\n",
    "fq_blank_line_end": b"@a\nACGT\n+\nIIII\n\n@b\nGG\n+\nJJ\n\n",
    "fa_record_longer_than_halo": b">long\n" + b"ACGTTGCA" * 4000 + b"\n>short\nAC\n",
    "fa_many_tiny_records": b"".join(b">%d\nA\n" % i for i in range(9000)),
}


@pytest.mark.parametrize("name", sorted(FALLBACK_INPUTS))
def test_fused_falls_back(lib, name):
    data = FALLBACK_INPUTS[name]
    opts = {"Reverse": True, "Complement": True}
    try:
        exp, exp_err = oracle.seq(data, opts), None
    except oracle.OracleError as e:
        exp, exp_err = None, str(e)
    try:
        (r, t), err = run(lib, data, opts), None
    except Exception as e:  # noqa: BLE001
        r, t, err = None, None, str(e)
    assert err == exp_err
    if exp is not None:
        assert r.data == exp[0] and list(r.elem_off) == exp[1]
        assert t["fused_blocks"] == 0


def test_fused_multi_block_partition(lib, monkeypatch):
    # bsk_run_buffer cuts the partition into record-aligned blocks; each block takes the tile path
    data = synth.fastq_reads(200 << 10, seed=14).tobytes()
    opts = {"Reverse": True, "Complement": True}
    exp = oracle.seq(data, opts)
    monkeypatch.setenv("BSK_BLOCK_BYTES", str(48 << 10))
    r, t = run(lib, data, opts)
    assert r.data == exp[0] and list(r.elem_off) == exp[1]
    assert t["fused_blocks"] >= 4
