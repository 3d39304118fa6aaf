"""rmdup: CUDA path vs the CPU oracle (first occurrence in input order wins)."""
import random

import pytest

import oracle
from bigseqkit_b200.api import BskError, Operator
from cases import EDGE_INPUTS, FQ_SIMPLE, fuzz_fasta, fuzz_fastq
from util import check_parity, run_lib, run_oracle

RMDUP_OPTS = [{}, {"BySeq": True}, {"ByName": True}, {"BySeq": True, "IgnoreCase": True}, {"IgnoreCase": True},
              {"BySeq": True, "OnlyPositiveStrand": True}, {"BySeq": True, "Config": {"LineWidth": 7}},
              {"Config": {"IDNCBI": True}}]


def dup_input(seed, fastq):
    rng = random.Random(seed)
    base = (fuzz_fastq(rng, n_rec=30, max_len=40) if fastq else fuzz_fasta(rng, n_rec=30, max_len=40, alphabet="ACGTacgt"))
    starts = oracle.frame(base)
    recs = [base[starts[i]:starts[i + 1]] for i in range(len(starts) - 1)]
    out = []
    for _ in range(120):
        r = rng.choice(recs)
        if rng.random() < 0.3:  # same subject, different case
            r = r.replace(b"A", b"a") if rng.random() < 0.5 else r
        out.append(r if r.endswith(b"\n") else r + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("opts", RMDUP_OPTS, ids=lambda o: str(o)[:60])
def test_rmdup_edge_inputs(lib, opts):
    for name, data in EDGE_INPUTS.items():
        check_parity(lib, "RmDup", data, opts)


@pytest.mark.parametrize("opts", RMDUP_OPTS[:5], ids=lambda o: str(o)[:60])
def test_rmdup_with_duplicates(lib, opts):
    for seed in range(3):
        for fastq in (False, True):
            data = dup_input(seed, fastq)
            got = check_parity(lib, "RmDup", data, opts)
            assert got is None or len(got[1]) - 1 < 120  # something was removed (None: both sides raised the same error)


def test_rmdup_kat(lib):
    # a/b share ACGT, key -861719356253734761 (xxh64("ACGT") as int64)
    data = b"@a\nACGT\n+\nIIII\n@b\nACGT\n+\nJJJJ\n@c\nAGGT\n+\nIIII\n"
    with Operator("RmDup", {"BySeq": True}, lib=lib) as o:
        r = o.call(data)
        assert r.data == b"@a\nACGT\n+\nIIII\n@c\nAGGT\n+\nIIII\n"
        assert o.rmdup_removed() == 1
        keys = o.rmdup_keys()
    assert keys[0] == keys[1] == -861719356253734761


def test_rmdup_keys_match_xxh64(lib):
    xxhash = pytest.importorskip("xxhash")
    rng = random.Random(5)
    data = fuzz_fasta(rng, n_rec=60, max_len=300, width=0)
    for opts in ({"BySeq": True}, {"ByName": True}, {}, {"BySeq": True, "IgnoreCase": True}):
        with Operator("RmDupPrepare", opts, lib=lib) as o:
            r = o.call(data)
            keys = o.rmdup_keys()
        assert keys == oracle.rmdup_keys(data, opts)
        assert r.n_records == len(keys)
    # independent pin: python-xxhash on the sequences
    with Operator("RmDupPrepare", {"BySeq": True}, lib=lib) as o:
        o.call(data)
        keys = o.rmdup_keys()
    starts = oracle.frame(data)
    for i, k in enumerate(keys):
        rec = data[starts[i]:starts[i + 1]]
        seq = b"".join(rec.split(b"\n")[1:])
        h = xxhash.xxh64(seq, seed=0).intdigest()
        assert k == (h - (1 << 64) if h >= 1 << 63 else h)


def test_rmdup_bad_flags(lib):
    check_parity(lib, "RmDup", FQ_SIMPLE, {"BySeq": True, "ByName": True})
    check_parity(lib, "RmDup", FQ_SIMPLE, {"OnlyPositiveStrand": True})


def test_rmdup_multi_block_partition(lib, monkeypatch):
    # the host entry point cuts a partition into record-aligned blocks; duplicates must be
    # recognised across blocks (history of fingerprints) and the output must not depend on the cut
    data = dup_input(11, True) * 3
    exp = oracle.rmdup(data, {"BySeq": True})
    monkeypatch.setenv("BSK_BLOCK_BYTES", "4096")
    with Operator("RmDup", {"BySeq": True}, lib=lib) as o:
        r = o.call(data)
        assert r.data == exp[0] and list(r.elem_off) == exp[1]
        assert o.rmdup_removed() == exp[2]


def test_rmdup_compacts_byte_ranges_like_the_formatter(lib, monkeypatch):
    # k_emit_contig (records printed as they stand) against k_emit (byte-wise formatter) and the oracle
    from bigseqkit_b200 import synth
    base = synth.fastq_reads(120 << 10, seed=81, dup_frac=0.3).tobytes()
    cases = {
        "plain": base,
        "no_final_newline": base[:-1],
        "last_record_is_a_duplicate": base + base[:base.index(b"\n@", 10) + 1],
        "fasta_single_line": synth.fasta_reads(1500, read_len=50, seed=82).tobytes() * 2,
    }
    for name, data in cases.items():
        exp = run_oracle("RmDup", data, {"BySeq": True})
        got = run_lib(lib, "RmDup", data, {"BySeq": True})
        assert got[0] == exp[0] and list(got[1]) == list(exp[1]), name
        monkeypatch.setenv("BSK_NO_CONTIG", "1")
        got2 = run_lib(lib, "RmDup", data, {"BySeq": True})
        monkeypatch.delenv("BSK_NO_CONTIG")
        assert got2[0] == exp[0] and list(got2[1]) == list(exp[1]), name


@pytest.mark.parametrize("opts", [{"BySeq": True}, {"ByName": True}, {}], ids=["by_seq", "by_name", "by_id"])
def test_rmdup_tile_front_end_matches_general_path(lib, monkeypatch, opts):
    # k_rmdup_tile (index + parse + hash in one pass) against the four-pass general path and the oracle
    from bigseqkit_b200 import synth
    from test_fused_path import _fixed_fastq
    reads = synth.fastq_reads(150 << 10, seed=83, dup_frac=0.3).tobytes()
    same_ids = b"".join(b"@id%d desc %d\nACGT%s\n+\nIIII%s\n" % (i % 97, i, b"A" * (100 + i % 50), b"I" * (100 + i % 50)) for i in range(1500))
    for name, data in (("reads", reads), ("repeated_ids", same_ids), ("rec128_tile_aligned", _fixed_fastq(1500, 8, 57, 84) * 2)):
        exp = run_oracle("RmDup", data, opts)
        with Operator("RmDup", opts, lib=lib) as o:
            r = o.call(data)
            t = o.timings()
            keys = o.rmdup_keys()
        assert r.data == exp[0] and list(r.elem_off) == list(exp[1]), (name, opts)
        assert t["fused_blocks"] == 1, (name, t)
        monkeypatch.setenv("BSK_NO_RMDUP_TILE", "1")
        with Operator("RmDup", opts, lib=lib) as o:
            r2 = o.call(data)
            assert o.timings()["fused_blocks"] == 0
            keys2 = o.rmdup_keys()
        monkeypatch.delenv("BSK_NO_RMDUP_TILE")
        assert r2.data == exp[0] and keys == keys2, (name, opts)
    # outside the tile grammar: "+name" lines, FASTA, --ignore-case -> general path, same answers
    for data, o2 in ((_fixed_fastq(300, 9, 80, 85, plus="x"), opts), (synth.fasta_reads(800, read_len=60, seed=86).tobytes() * 2, opts),
                     (reads, dict(opts, IgnoreCase=True))):
        exp = run_oracle("RmDup", data, o2)
        with Operator("RmDup", o2, lib=lib) as o:
            r = o.call(data)
            assert o.timings()["fused_blocks"] == 0
        assert r.data == exp[0]


DUP_FILE_OPTS = [{"BySeq": True}, {"ByName": True}, {}, {"BySeq": True, "IgnoreCase": True},
                 {"BySeq": True, "Config": {"LineWidth": 7}}, {"Config": {"IDNCBI": True}}]


def _dup_files(lib, data, opts):
    o = dict(opts, DupSeqsFile="dups.fx", DupNumFile="dups.txt")
    with Operator("RmDup", o, lib=lib) as op:
        r = op.call(data)
        return r.data, op.rmdup_dup_seqs(), op.rmdup_dup_num(), op.rmdup_removed()


def test_rmdup_dup_files_kat(lib):
    # hand-derived from bigseqkit-lib/rmdup.go:180-239: a/c/d share ACGT, b/e share TTTT
    d = (b"@a x\nACGT\n+\nIIII\n@b\nTTTT\n+\nIIII\n@c y\nACGT\n+\nJJJJ\n@d\nACGT\n+\nKKKK\n@e\nTTTT\n+\nIIII\n"
         b"@f\nGG\n+\nII\n")
    kept, seqs, num, removed = _dup_files(lib, d, {"BySeq": True})
    assert kept == b"@a x\nACGT\n+\nIIII\n@b\nTTTT\n+\nIIII\n@f\nGG\n+\nII\n"
    assert seqs == b"@c y\nACGT\n+\nJJJJ\n@d\nACGT\n+\nKKKK\n@e\nTTTT\n+\nIIII\n"
    assert num == b"3\ta, c, d\n2\tb, e\n"
    assert removed == 3
    assert oracle.rmdup_dups(d, {"BySeq": True}) == (seqs, num)
    # nothing removed: both empty
    assert _dup_files(lib, d, {})[1:3] == (b"", b"")


@pytest.mark.parametrize("opts", DUP_FILE_OPTS, ids=lambda o: str(o)[:60])
def test_rmdup_dup_files_match_oracle(lib, opts, monkeypatch):
    from bigseqkit_b200 import synth
    inputs = [dup_input(21, False), dup_input(22, True), synth.fastq_reads(60 << 10, seed=87, dup_frac=0.3).tobytes()]
    for data in inputs:
        exp_kept = oracle.rmdup(data, opts)
        exp = oracle.rmdup_dups(data, opts)
        kept, seqs, num, removed = _dup_files(lib, data, opts)
        assert kept == exp_kept[0] and removed == exp_kept[2]
        assert (seqs, num) == exp
        # removed + kept records together are the input's records
        assert len(oracle.frame(seqs)) - 1 + len(exp_kept[1]) - 1 == len(oracle.frame(data)) - 1
    # several blocks per partition: the first member of a group sits in an earlier block
    monkeypatch.setenv("BSK_BLOCK_BYTES", "4096")
    for data in inputs[:2]:
        exp = oracle.rmdup_dups(data * 3, opts)
        assert _dup_files(lib, data * 3, opts)[1:3] == exp


def test_rmdup_dup_num_rejected_on_the_sharded_path(lib):
    with Operator("RmDup", {"BySeq": True, "DupNumFile": "x"}, lib=lib) as op:
        with pytest.raises(BskError):
            op.rmdup_prepare_device(0, 0, 0, 0)
