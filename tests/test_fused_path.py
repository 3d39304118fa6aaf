"""The single-pass tile kernels for short records (k_fastq_inplace.cu; the general path behind them): parity with the
oracle on inputs that span many tiles, and clean fall-back to the general path for everything the tile scheme cannot
represent."""
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
def test_parity_many_tiles(lib, opts):
    for name, data in big_inputs().items():
        try:
            exp = oracle.seq(data, opts)
        except oracle.OracleError:
            continue  # e.g. -q on FASTA: covered by the general-path tests
        r, t = run(lib, data, opts)
        assert r.data == exp[0], (name, opts)
        assert list(r.elem_off) == exp[1], (name, opts)


def test_plain_formatting_takes_the_general_path(lib):
    data = synth.fasta_reads(2500, read_len=100, seed=3).tobytes()
    r, t = run(lib, data, {})
    exp = oracle.seq(data, {})
    assert r.data == exp[0] and list(r.elem_off) == exp[1] and t["fused_blocks"] == 0


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


# ---------------------------------------------------------------- same-layout FASTQ kernel (k_fastq_inplace.cu)
def _fixed_fastq(n_rec, hdr_len, seq_len, seed, alphabet="ACGTN", plus=""):
    rng = random.Random(seed)
    out = []
    for i in range(n_rec):
        h = ("%0*d" % (hdr_len, i))[-hdr_len:] if hdr_len else ""
        s = "".join(rng.choice(alphabet) for _ in range(seq_len))
        q = "".join(chr(rng.randint(33, 74)) for _ in range(seq_len))
        if plus and q[:1] == "@":
            q = "I" + q[1:]  # "\n@" after a "+name" line would be framed as a record start
        out.append("@%s\n%s\n+%s\n%s\n" % (h, s, plus, q))
    return "".join(out).encode()


INPLACE_OPTS = [
    {"Reverse": True, "Complement": True},
    {"Reverse": True},
    {"Complement": True},
    {},
    {"Name": True, "Seq": True, "Reverse": True, "Complement": True, "LowerCase": True},
    {"Dna2rna": True, "Reverse": True},
]


def inplace_inputs():
    return {
        # 64-byte records: every 12 / 16 / 20 KiB tile boundary is a record start
        "rec64_tile_aligned": _fixed_fastq(2000, 8, 25, 21),
        # records of 1 KiB: tile boundaries fall at every phase of a record, starts on boundaries included
        "rec1024": _fixed_fastq(150, 10, 504, 22),
        # 48-byte records: > 2000 lines per region, several passes of the CTA over the line list
        "rec48_many_lines": _fixed_fastq(3000, 6, 18, 30),
        "reads150": synth.fastq_reads(400 << 10, seed=23).tobytes(),
        "iupac_lower": _fixed_fastq(700, 12, 151, 24, alphabet="ACGTNacgtnRYKMSWBDHVrykm"),
        "no_final_newline": synth.fastq_reads(90 << 10, seed=25).tobytes()[:-1],
        "len250": _fixed_fastq(400, 30, 250, 26),        # 63-64 words per segment: register path at its limit
        "len251_to_600": b"".join(_fixed_fastq(1, 20, 251 + 7 * i, 100 + i) for i in range(50)),  # byte-pair path
        "empty_seq_and_header": b"@\n\n+\n\n@a\nA\n+\nI\n" * 250,
        "qual_starts_with_at_and_plus": b"@r\nACGT\n+\n@+@+\n@s\nGG\n+\n+@\n" * 250,
        # short first record (small first-attempt scan halo), then records that end far into the halo: rescan path
        "short_then_long": _fixed_fastq(1, 4, 20, 27) + _fixed_fastq(120, 10, 900, 28) + _fixed_fastq(40, 10, 1700, 29),
        # long first record, then many short ones: more elements than the first record suggests (offset array re-run)
        "long_then_short": _fixed_fastq(1, 10, 900, 44) + _fixed_fastq(4000, 6, 40, 45),
        "tiny_file": b"@a\nACGT\n+\nIIII\n",
        "tiny_file_no_nl": b"@a\nAC\n+\nII",
    }


@pytest.mark.parametrize("opts", INPLACE_OPTS, ids=lambda o: str(o)[:50])
def test_inplace_parity(lib, opts):
    for name, data in inplace_inputs().items():
        exp = oracle.seq(data, opts)
        r, t = run(lib, data, opts)
        assert r.data == exp[0], (name, opts)
        assert list(r.elem_off) == exp[1], (name, opts)
        # main kernel + offset expansion (+ one re-run of the expansion when the first record under-estimates the count)
        assert t["fused_blocks"] == 1 and t["kernel_launches"] == (3 if name == "long_then_short" else 2), (name, opts, t)


INPLACE_OUTSIDE_GRAMMAR = {
    "plus_with_name": _fixed_fastq(300, 9, 80, 31, plus="x y"),
    "one_plus_with_name_late": _fixed_fastq(900, 9, 80, 32) + b"@z\nAC\n+z\nII\n" + _fixed_fastq(50, 9, 80, 33),
    "record_longer_than_halo": _fixed_fastq(3, 5, 6000, 34) + _fixed_fastq(200, 9, 80, 35),
    "multiline_in_the_middle": _fixed_fastq(400, 9, 80, 36) + b"@m\nACGT\nAC\n+\nIIII\nII\n" + _fixed_fastq(10, 9, 80, 37),
    "seq_line_starts_with_plus": _fixed_fastq(100, 9, 80, 38) + b"@p\n+CGT\n+\nIIII\n",
    "missing_first_marker": b"ACGT\n+\nIIII\n" + _fixed_fastq(100, 9, 80, 39),
    "trailing_blank_line": _fixed_fastq(100, 9, 80, 40) + b"\n",
}


@pytest.mark.parametrize("name", sorted(INPLACE_OUTSIDE_GRAMMAR))
def test_inplace_leaves_other_grammars_to_the_general_paths(lib, name):
    data = INPLACE_OUTSIDE_GRAMMAR[name]
    opts = {"Reverse": True, "Complement": True}
    try:
        exp, exp_err = oracle.seq(data, opts), None
    except oracle.OracleError as e:
        exp, exp_err = None, str(e)
    try:
        (r, t), err = run(lib, data, opts), None
    except Exception as e:  # noqa: BLE001
        r, t, err = None, None, str(e)
    assert err == exp_err, name
    if exp is not None:
        assert r.data == exp[0] and list(r.elem_off) == exp[1], name


def test_records_that_start_on_a_tile_boundary(lib):
    # 64-byte records: every tile boundary (20 KiB tiles here, 16 KiB tiles of the FASTA index) is a record start
    data = _fixed_fastq(2000, 8, 25, 41)
    for opts in ({"Reverse": True, "Complement": True}, {"MinLen": 5}, {"Name": True}):
        exp = oracle.seq(data, opts)
        r, t = run(lib, data, opts)
        assert r.data == exp[0] and list(r.elem_off) == exp[1]
    fa = b"".join(b">%05d\n%s\n" % (i, b"ACGTACGTAC" * 5 + b"ACGTAC") for i in range(2000))  # 64-byte FASTA records
    exp = oracle.seq(fa, {"Complement": True})
    r, t = run(lib, fa, {"Complement": True})
    assert r.data == exp[0] and list(r.elem_off) == exp[1]


@pytest.mark.parametrize("shape", ["1", "2", "3", "4"])
def test_inplace_cta_shapes(lib, monkeypatch, shape):
    # the CTA shapes kept for A/B runs (BSK_FQ_SHAPE: 512 x 3, 256 x 8, 384 x 5, 1024 x 2) give the same bytes as the default
    if lib.path.endswith("libbsk_emu.so") and shape == "3":
        pytest.skip("covered on the GPU")
    monkeypatch.setenv("BSK_FQ_SHAPE", shape)
    opts = {"Reverse": True, "Complement": True}
    names = ("reads150", "len250", "rec64_tile_aligned", "short_then_long", "no_final_newline", "rec48_many_lines", "tiny_file")
    if lib.path.endswith("libbsk_emu.so"):
        names = ("no_final_newline", "rec64_tile_aligned", "tiny_file")  # the host emulator is slow (OS threads)
    for name in names:
        data = inplace_inputs()[name]
        exp = oracle.seq(data, opts)
        r, t = run(lib, data, opts)
        assert r.data == exp[0] and list(r.elem_off) == exp[1], (name, shape)


@pytest.mark.parametrize("group", ["8", "32"])
@pytest.mark.parametrize("opts", [{"Reverse": True, "Complement": True}, {"Reverse": True}, {"Complement": True}], ids=str)
def test_inplace_lane_groups(lib, monkeypatch, group, opts):
    # both lane groupings of the in-place transform on every input (the library picks one from the first record)
    monkeypatch.setenv("BSK_FQ_GROUP", group)
    for name in ("reads150", "len250", "len251_to_600", "rec64_tile_aligned", "empty_seq_and_header", "short_then_long",
                 "no_final_newline", "rec48_many_lines"):
        data = inplace_inputs()[name]
        exp = oracle.seq(data, opts)
        r, t = run(lib, data, opts)
        assert r.data == exp[0] and list(r.elem_off) == exp[1], (name, group)
        assert t["fused_blocks"] == 1 and t["kernel_launches"] == 2
