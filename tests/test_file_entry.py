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
