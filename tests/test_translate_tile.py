"""translate on FASTA wrapped at one width, sequences read in place (k_fasta_tile.cu + k_translate_tile.cu) against the
oracle (Translate.Call, bigseqkit-lib/translate.go:66-145).  Inputs are shaped to hit the seams: records longer than a
tile, wrap widths around 48 (the fast path's limit) and around the 16-byte output window, ambiguous / lower-case bases
and gaps (careful path), ragged wrapping and a missing final newline (the block goes to the general path)."""
import random

import numpy as np
import pytest

import oracle
from bigseqkit_b200 import synth
from bigseqkit_b200.api import BskError, Operator


def _wrap(seq, w):
    if w <= 0:
        return seq + b"\n"
    return b"".join(seq[i:i + w] + b"\n" for i in range(0, len(seq), w)) or b"\n"


def _fasta(rng, lens, width, alphabet=b"ACGT", sprinkle=b"", frac=0.0):
    out = []
    for i, L in enumerate(lens):
        s = bytearray(rng.choice(alphabet) for _ in range(L))
        for _ in range(int(L * frac)):
            s[rng.randrange(L)] = rng.choice(sprinkle)
        out.append(b">r%d some description %d\n" % (i, L) + _wrap(bytes(s), width))
    return b"".join(out)


def _check(lib, data, opts, expect_tile=True):
    try:
        exp, exp_off = oracle.translate(data, opts)
    except oracle.OracleError as e:
        with Operator("Translate", opts, lib=lib) as op:
            with pytest.raises(BskError) as ei:
                op.call(data)
        assert str(ei.value) == str(e)
        return
    with Operator("Translate", opts, lib=lib) as op:
        got = op.call(data)
        fused = op.timings()["fused_blocks"]
    assert got.data == exp
    assert list(got.elem_off) == exp_off
    if expect_tile is not None:
        assert (fused > 0) == expect_tile


@pytest.mark.parametrize("width", [60, 0, 48, 49, 47, 100, 7])
def test_widths(lib, width):
    rng = random.Random(width)
    data = _fasta(rng, [3000, 300, 26001, 9, 3, 1500, 64, 48, 47, 46], width)
    _check(lib, data, {"Frame": ["6"]})


@pytest.mark.parametrize("frames", [["1"], ["2"], ["3"], ["-1"], ["-2"], ["-3"], ["1", "-1"], ["3", "2", "1"], ["6"]])
def test_frames(lib, frames):
    rng = random.Random(5)
    data = _fasta(rng, [999, 1000, 1001, 30, 31, 32, 33, 5000], 60)
    _check(lib, data, {"Frame": frames})


@pytest.mark.parametrize("lw", [60, 0, 16, 15, 17, 100, 1])
def test_output_widths(lib, lw):
    rng = random.Random(100 + lw)
    data = _fasta(rng, [2000, 333, 48, 51, 96, 99], 60)
    _check(lib, data, {"Frame": ["6"], "Config": {"LineWidth": lw}})


@pytest.mark.parametrize("opts", [{"TranslTable": 2}, {"TranslTable": 11, "InitCodonAsM": True}, {"Clean": True}, {"AllowUnknownCodon": True},
                                  {"Clean": True, "InitCodonAsM": True, "TranslTable": 4}], ids=str)
def test_options(lib, opts):
    rng = random.Random(11)
    data = b">atg\nATGTTGCTGTAAATTGTG\n" + _fasta(rng, [1200, 600, 3000], 60)
    _check(lib, data, dict(opts, Frame=["6"]))


def test_careful_path_bases(lib):
    rng = random.Random(12)
    data = _fasta(rng, [2400, 1200], 60, sprinkle=b"NRYKMnacgt-", frac=0.03)
    _check(lib, data, {"Frame": ["6"]})
    data = _fasta(rng, [900], 60, alphabet=b"ACGU")          # RNA: nothing is "plain ACGT"
    _check(lib, data, {"Frame": ["6"]})
    data = _fasta(rng, [900], 60, alphabet=b"acgt")
    _check(lib, data, {"Frame": ["1", "-1"]})


def test_unknown_codon_error(lib):
    rng = random.Random(13)
    data = _fasta(rng, [600, 600], 60) + b">bad\nACG?ACGTACGT\n" + _fasta(rng, [300], 60)
    _check(lib, data, {"Frame": ["1"]})
    _check(lib, data, {"Frame": ["1"], "AllowUnknownCodon": True})


def test_too_short_error(lib):
    _check(lib, b">a\nACGTACGT\n>b\nAC\n>c\nACGTAC\n", {"Frame": ["1"]})


def test_declines(lib):
    rng = random.Random(14)
    body = bytes(rng.choice(b"ACGT") for _ in range(600))
    ragged = b">a\n" + body[:60] + b"\n" + body[60:100] + b"\n" + body[100:160] + b"\n>b\n" + _wrap(body, 60)
    _check(lib, ragged, {"Frame": ["6"]}, expect_tile=False)
    _check(lib, b">a\n" + _wrap(body, 60)[:-1], {"Frame": ["6"]}, expect_tile=False)   # no final newline
    _check(lib, b">a\n" + _wrap(body, 60), {"Frame": ["6"], "Trim": True}, expect_tile=False)
    _check(lib, b">a\n" + _wrap(body, 60), {"Frame": ["6"], "AppendFrame": True}, expect_tile=False)
    fq = synth.fastq_reads(4096, seed=3).tobytes()
    _check(lib, fq, {"Frame": ["1"]}, expect_tile=False)


def test_blank_line_after_last_full_line(lib):
    rng = random.Random(15)
    body = bytes(rng.choice(b"ACGT") for _ in range(120))
    _check(lib, b">a\n" + body[:60] + b"\n" + body[60:] + b"\n\n>b\n" + body[:30] + b"\n", {"Frame": ["6"]}, expect_tile=None)


@pytest.mark.gpu
def test_native_cds_block():
    arr, _ = synth.native_cds(48 << 20, seed=45)
    opts = {"Frame": ["6"]}
    exp = oracle.run_mt_full("translate", arr.ctypes.data, arr.nbytes, opts, 8)
    with Operator("Translate", opts, device=0) as op:
        got = op.call((arr.ctypes.data, arr.nbytes))
        assert op.timings()["fused_blocks"] > 0
    assert got.data == exp["data"].tobytes()
    assert np.array_equal(np.asarray(got.elem_off, dtype=np.uint64), exp["elem_off"])
