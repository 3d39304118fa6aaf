"""File-level entry points of the C ABI: record-aligned shard planning and range -> file runs (ordered merged output)."""
import oracle
from bigseqkit_b200 import api, synth
from bigseqkit_b200.api import Operator


def test_shard_bounds_and_merged_output(lib, tmp_path):
    data = synth.fastq_reads(300 << 10, seed=95).tobytes()
    src = tmp_path / "reads.fq"
    src.write_bytes(data)
    starts = set(oracle.frame(data))
    for shards in (1, 2, 3, 7):
        b = api.shard_bounds(str(src), shards, lib=lib)
        assert b[0] == 0 and b[-1] == len(data) and len(b) == shards + 1
        assert all(x in starts for x in b) and b == sorted(b)
    # every shard is run separately and written at its offset of one merged file (same bytes as a single run)
    opts = {"Reverse": True, "Complement": True}
    exp, _ = oracle.seq(data, opts)
    b = api.shard_bounds(str(src), 3, lib=lib)
    out = tmp_path / "out.fq"
    off = 0
    with Operator("SeqTransform", opts, lib=lib) as op:
        for r in range(3):
            ob, nr, ne = op.call_file(str(src), b[r], b[r + 1] - b[r], str(out), off, partition_id=r)
            assert nr == ne and ob > 0
            off += ob
    assert out.read_bytes() == exp
    # stats over a file range needs no output file
    with Operator("Stats", {"Tabular": True}, lib=lib) as op:
        ob, nr, ne = op.call_file(str(src))
        assert ob == 0 and nr == len(oracle.frame(data)) - 1
        assert op.stats_render() == oracle.stats(data, {"Tabular": True})[1]


def test_shard_bounds_fasta_and_tricky_quality_lines(lib, tmp_path):
    fa = synth.fasta_cds(100 << 10, seed=96).tobytes()
    p = tmp_path / "cds.fa"
    p.write_bytes(fa)
    starts = set(oracle.frame(fa))
    assert all(x in starts for x in api.shard_bounds(str(p), 5, lib=lib))
    tricky = b"@a\nACGT\n+\n@III\n" * 4000  # quality lines start with '@' right after a bare '+'
    q = tmp_path / "t.fq"
    q.write_bytes(tricky)
    assert all(x % 15 == 0 for x in api.shard_bounds(str(q), 9, lib=lib))


def test_streamed_range_many_blocks(lib, tmp_path, monkeypatch):
    # bsk_run_file streams the range through a ring of pinned slots: many small blocks, state carried from block to block,
    # one record longer than a block (the slot grows), output written block by block at its running offset
    monkeypatch.setenv("BSK_BLOCK_BYTES", "8192")
    fq = synth.fastq_reads(150 << 10, seed=97, dup_frac=0.3).tobytes()
    fa = (synth.fasta_cds(60 << 10, seed=98).tobytes() + b">long\n" + b"ACGTTGCAAC" * 3000 + b"\n"
          + synth.fasta_cds(20 << 10, seed=99).tobytes())
    cases = [
        ("SeqTransform", {"Reverse": True, "Complement": True}, fq, oracle.seq),
        ("SeqTransform", {"MinLen": 120}, fq, oracle.seq),
        ("RmDup", {"BySeq": True}, fq, lambda d, o: oracle.rmdup(d, o)[:2]),
        ("Translate", {"Frame": ["6"]}, fa, oracle.translate),
        ("SeqTransform", {"Complement": True}, fa, oracle.seq),
        ("Fq2Fa", {}, fq, oracle.fq2fa),
    ]
    for k, (opn, opts, data, ref) in enumerate(cases):
        src = tmp_path / ("in%d" % k)
        src.write_bytes(data)
        out = tmp_path / ("out%d" % k)
        exp = ref(data, opts)[0]
        with Operator(opn, opts, lib=lib) as op:
            ob, nr, ne = op.call_file(str(src), 0, 0, str(out), 0)
        assert ob == len(exp), (opn, opts)
        assert out.read_bytes() == exp, (opn, opts)
    # stats: no output, the histogram accumulates over the blocks
    src = tmp_path / "s.fq"
    src.write_bytes(fq)
    with Operator("Stats", {"Tabular": True, "All": True}, lib=lib) as op:
        ob, nr, ne = op.call_file(str(src))
        assert ob == 0 and op.stats_render() == oracle.stats(fq, {"Tabular": True, "All": True})[1]
    # empty file
    e = tmp_path / "empty.fq"
    e.write_bytes(b"")
    with Operator("SeqTransform", {}, lib=lib) as op:
        assert op.call_file(str(e), 0, 0, str(tmp_path / "eo"), 0)[0] == 0
