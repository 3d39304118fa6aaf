"""Small runs of every tile kernel + the window formatter for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bigseqkit_b200 import Operator, synth
fq = synth.fastq_reads(192 << 10, seed=7, dup_frac=0.2).tobytes()
fa = synth.fasta_cds(128 << 10, seed=8).tobytes()
checks = [("SeqTransform", {"Reverse": True, "Complement": True}, fq[:-1], oracle.seq),
          ("SeqTransform", {"Reverse": True, "Complement": True}, fa, oracle.seq),
          ("SeqTransform", {"MinLen": 100, "Reverse": True}, fq, oracle.seq),
          ("RmDup", {"BySeq": True}, fq, lambda d, o: oracle.rmdup(d, o)[:2]),
          ("Translate", {"Frame": ["6"]}, fa, oracle.translate),
          ("Translate", {"Frame": ["2", "-3"], "Clean": True, "InitCodonAsM": True, "AllowUnknownCodon": True},
           fa.replace(b"ACG", b"A-N", 50), oracle.translate),
          ("Fq2Fa", {}, fq, oracle.fq2fa),
          ("SubseqTransform", {"Region": "5:-5"}, fq, oracle.subseq)]
for name, opts, data, fn in checks:
    with Operator(name, opts, device=0) as op:
        r = op.call(data)
    exp = fn(data, opts)
    print(name, opts, "ok" if r.data == exp[0] else "MISMATCH", flush=True)
for opts, data in (({"Tabular": True, "All": True}, fq), ({"Tabular": True}, fa)):
    with Operator("Stats", opts, device=0) as op:
        op.call(data)
        print("Stats", opts, "ok" if op.stats_render() == oracle.stats(data, opts)[1] else "MISMATCH", flush=True)
# locate tile kernel (equal-length ACGT panel on wrapped contigs), raw-element operators, the streamed file path
ctg, _ = synth.native_contigs(512 << 10, seed=9, max_len=200_000)
ctg = ctg.tobytes()
lopts = {"Pattern": synth.pattern_panel(200, 12, 40)}
with Operator("Locate", lopts, device=0) as op:
    r = op.call(ctg)
    print("Locate", "ok" if r.data == oracle.locate(ctg, lopts)[0] and op.timings()["fused_blocks"] > 0 else "MISMATCH", flush=True)
with Operator("Duplicate", {"Times": 3}, device=0) as op:
    print("Duplicate", "ok" if op.call(fq).data == oracle.duplicate(fq, 3)[0] else "MISMATCH", flush=True)
with Operator("Range", {"Start": 5, "End": 400}, device=0) as op:
    print("Range", "ok" if op.call(fa).data == oracle.range_(fa, 5, 400)[0] else "MISMATCH", flush=True)
import tempfile
with tempfile.TemporaryDirectory() as td:
    os.environ["BSK_BLOCK_BYTES"] = "32768"
    src, dst = os.path.join(td, "in.fq"), os.path.join(td, "out.fq")
    open(src, "wb").write(fq)
    with Operator("SeqTransform", {"Reverse": True, "Complement": True}, device=0) as op:
        op.call_file(src, 0, 0, dst, 0)
    ok = open(dst, "rb").read() == oracle.seq(fq, {"Reverse": True, "Complement": True})[0]
    print("bsk_run_file (streamed, 32 KiB blocks)", "ok" if ok else "MISMATCH", flush=True)
    del os.environ["BSK_BLOCK_BYTES"]
with Operator("RmDup", {"BySeq": True, "DupSeqsFile": "d", "DupNumFile": "D"}, device=0) as op:
    op.call(fq)
    got = (op.rmdup_dup_seqs(), op.rmdup_dup_num())
    print("RmDup -d -D", "ok" if got == oracle.rmdup_dups(fq, {"BySeq": True}) else "MISMATCH", flush=True)
