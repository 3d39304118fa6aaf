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
with Operator("RmDup", {"BySeq": True, "DupSeqsFile": "d", "DupNumFile": "D"}, device=0) as op:
    op.call(fq)
    got = (op.rmdup_dup_seqs(), op.rmdup_dup_num())
    print("RmDup -d -D", "ok" if got == oracle.rmdup_dups(fq, {"BySeq": True}) else "MISMATCH", flush=True)
