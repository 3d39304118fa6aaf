"""locate with an equal-length ACGT panel on FASTA: the single-pass tile kernel (k_locate_tile.cu) against the oracle
(Locate.Call default exact path, bigseqkit-lib/locate.go:395-769).  Inputs are shaped to hit the kernel's seams: contigs
longer than a tile, wrapped at widths around the lane span, header lines across tile boundaries, blank lines, invalid
bases, a missing final newline."""
import random

import numpy as np
import pytest

import oracle
from bigseqkit_b200 import synth
from bigseqkit_b200.api import Operator

TILE = 23520


def _wrap(seq, w):
    if w <= 0:
        return seq + b"\n"
    return b"".join(seq[i:i + w] + b"\n" for i in range(0, len(seq), w)) or b"\n"


def _fasta(rng, lens, width, n_frac=0.0, lower_frac=0.0, hdr_pad=0):
    out = []
    for i, L in enumerate(lens):
        s = bytearray(rng.choice(b"ACGT") for _ in range(L))
        for _ in range(int(L * n_frac)):
            s[rng.randrange(L)] = ord("N")
        for _ in range(int(L * lower_frac)):
            j = rng.randrange(L)
            s[j] = s[j] | 0x20
        out.append(b">c%d %s\n" % (i, b"x" * hdr_pad) + _wrap(bytes(s), width))
    return b"".join(out)


def _panel(rng, data, k, n, from_data=0.5):
    seqs = b"".join(l for l in data.split(b"\n") if not l.startswith(b">"))
    pats = []
    while len(pats) < n:
        if rng.random() < from_data and len(seqs) > k:
            j = rng.randrange(len(seqs) - k)
            p = seqs[j:j + k].upper()
            if set(p) <= set(b"ACGT"):
                pats.append(p.decode())
        else:
            pats.append("".join(rng.choice("ACGT") for _ in range(k)))
    return pats


def _check(lib, data, opts, expect_tile=True):
    exp, exp_off = oracle.locate(data, opts)
    with Operator("Locate", opts, lib=lib) as op:
        got = op.call(data)
        fused = op.timings()["fused_blocks"]
    assert got.data == exp
    assert list(got.elem_off) == exp_off
    if expect_tile is not None:
        assert (fused > 0) == expect_tile


@pytest.mark.parametrize("width", [60, 0, 7, 95, 96, 97, 1])
def test_contigs_across_tiles(lib, width):
    rng = random.Random(1000 + width)
    data = _fasta(rng, [30000, 5, 0, 41000, 12, 977], width, n_frac=0.002)
    pats = _panel(rng, data, 12, 40)
    # width 1: more newlines than bases around every lane boundary -> the kernel declines, the general path answers
    _check(lib, data, {"Pattern": pats}, expect_tile=None if width == 1 else True)


@pytest.mark.parametrize("k", [2, 5, 12, 16])
def test_pattern_lengths(lib, k):
    rng = random.Random(k)
    data = _fasta(rng, [26000, 300, 9000], 60)
    _check(lib, data, {"Pattern": _panel(rng, data, k, 8 if k < 6 else 60)})


def test_header_across_tile_boundary(lib):
    rng = random.Random(7)
    first = TILE - 40  # the second header starts a few bytes before the end of tile 0 and ends in tile 1
    body = _wrap(bytes(rng.choice(b"ACGT") for _ in range(first * 60 // 61 - 30)), 60)
    for pad in (0, 20, 200, 900):
        data = b">a\n" + body + b">b " + b"ACGTACGTACGT" * (pad // 12 + 1) + b"\n" + _wrap(bytes(rng.choice(b"ACGT") for _ in range(5000)), 60)
        pats = _panel(rng, data, 12, 30) + ["ACGTACGTACGT"]
        _check(lib, data, {"Pattern": pats})


def test_long_header_declines_to_general_path(lib):
    rng = random.Random(8)
    data = b">a " + b"ACGT" * 400 + b"\n" + _wrap(bytes(rng.choice(b"ACGT") for _ in range(3000)), 60)
    _check(lib, data, {"Pattern": ["ACGTACGTACGT", "GGGGGGGGGGGG"]}, expect_tile=False)


@pytest.mark.parametrize("opts", [{"OnlyPositiveStrand": True}, {"IgnoreCase": True}, {"NonGreedy": True}, {"Bed": True}, {"Gtf": True},
                                  {"HideMatched": True}, {"Config": {"IDNCBI": True}}], ids=str)
def test_options(lib, opts):
    rng = random.Random(9)
    data = _fasta(rng, [25000, 8000], 60, lower_frac=0.3 if opts.get("IgnoreCase") else 0.0)
    pats = _panel(rng, data, 6, 12) + ["AAAAAA", "ACACAC"]
    _check(lib, data, dict(opts, Pattern=pats))


def test_lowercase_sequence_is_case_sensitive(lib):
    rng = random.Random(10)
    data = _fasta(rng, [24000], 60, lower_frac=0.2)
    _check(lib, data, {"Pattern": _panel(rng, data, 8, 50)})
    _check(lib, data, {"Pattern": [p.lower() for p in _panel(rng, data, 4, 10)]})


def test_edges(lib):
    rng = random.Random(11)
    body = bytes(rng.choice(b"ACGT") for _ in range(2000))
    pats = _panel(rng, b">x\n" + body + b"\n", 12, 20)
    _check(lib, b">x\n" + body, {"Pattern": pats})                         # no final newline
    _check(lib, b">x\n" + body[:600] + b"\n\n\n" + body[600:] + b"\n", {"Pattern": pats})  # blank lines inside a record
    _check(lib, b">x\n\n>y\n" + body + b"\n>z", {"Pattern": pats})         # empty records, header-only tail
    _check(lib, b">only\n", {"Pattern": pats})
    _check(lib, b">p\nACGTACGTAC\n", {"Pattern": ["ACGTACGTACGT"]})        # pattern longer than the sequence
    _check(lib, b">dup\n" + body + b"\n", {"Pattern": [pats[0], pats[0], pats[1]]})  # the same pattern twice
    _check(lib, b">pal\nACGTACGTACGT\n", {"Pattern": ["ACGT"]})            # palindrome: '+' and '-' needles share a code


def test_large_panel_takes_general_path(lib):
    """more than 2048 needle codes (patterns x strands) would overfill the fingerprint table of the tile kernel"""
    rng = random.Random(21)
    data = _fasta(rng, [6000, 300], 60)
    pats = sorted(set(_panel(rng, data, 12, 1100, from_data=0.05)))
    _check(lib, data, {"Pattern": pats}, expect_tile=len(pats) * 2 <= 2048)
    _check(lib, data, {"Pattern": pats[:1000]}, expect_tile=True)


def test_mixed_panel_takes_general_path(lib):
    rng = random.Random(12)
    data = _fasta(rng, [3000], 60)
    _check(lib, data, {"Pattern": ["ACGT", "ACGTA"]}, expect_tile=False)
    _check(lib, data, {"Pattern": ["ACGN"]}, expect_tile=False)
    _check(lib, data, {"Pattern": ["ACGT"], "Circular": True}, expect_tile=False)


@pytest.mark.gpu
def test_native_contigs_block():
    arr, _ = synth.native_contigs(24 << 20, seed=44, max_len=2_000_000)
    opts = {"Pattern": synth.pattern_panel(1000, 12, 40)}
    exp = oracle.run_mt_full("locate", arr.ctypes.data, arr.nbytes, opts, 8)
    with Operator("Locate", opts, device=0) as op:
        got = op.call((arr.ctypes.data, arr.nbytes))
        assert op.timings()["fused_blocks"] > 0
    assert got.data == exp["data"].tobytes()
    assert np.array_equal(np.asarray(got.elem_off, dtype=np.uint64), exp["elem_off"])
