#!/usr/bin/env python
"""Device-resident timings of the other operators on the BASELINE.json config shapes (one JSON line per operator).
Not the driver's bench (that is bench.py, configs[1]); these lines document stats / rmdup / translate / locate."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ops", default="stats_fasta,stats_fastq,rmdup,translate,locate,subseq,grep_id,fq2fa")
    ap.add_argument("--cpu", action="store_true", help="also time the oracle port on all host cores (seq/stats/rmdup)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import oracle
    from bigseqkit_b200 import Operator, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    nbytes = args.mib << 20
    rng = np.random.Generator(np.random.PCG64(40))
    panel = ["".join("ACGT"[i] for i in rng.integers(0, 4, 12)) for _ in range(1000)]
    cases = {
        "stats_fasta": ("Stats", {"Tabular": True}, lambda: synth.fasta_reads(nbytes // 114, read_len=100, seed=1), "stats", 1.0),
        "stats_fastq": ("Stats", {"Tabular": True, "All": True}, lambda: synth.fastq_reads(nbytes, seed=2), "stats", 1.0),
        "rmdup": ("RmDup", {"BySeq": True}, lambda: synth.fastq_reads(nbytes, seed=3, dup_frac=0.2), "rmdup", None),
        "translate": ("Translate", {"Frame": ["6"]}, lambda: synth.fasta_cds(nbytes, seed=5), None, None),
        "seq_fasta_rc": ("SeqTransform", {"Reverse": True, "Complement": True}, lambda: synth.fasta_cds(nbytes, seed=5), None, None),
        "seq_fastq_minlen": ("SeqTransform", {"MinLen": 100, "Reverse": True, "Complement": True},
                             lambda: synth.fastq_reads(nbytes, seed=2), None, None),
        "seq_fasta_reads": ("SeqTransform", {}, lambda: synth.fasta_reads(nbytes // 114, read_len=100, seed=1), None, None),
        "subseq": ("SubseqTransform", {"Region": "10:-10"}, lambda: synth.fastq_reads(nbytes, seed=2), None, None),
        "fq2fa": ("Fq2Fa", {}, lambda: synth.fastq_reads(nbytes, seed=2), None, None),
        "grep_id": ("Grep", {"Pattern": ["SIM:1:FC:3:2208:1391:14437"], "InvertMatch": True},
                    lambda: synth.fastq_reads(nbytes, seed=2), None, None),
        "locate": ("Locate", {"Pattern": panel}, lambda: synth.fasta_contigs(min(nbytes, 256 << 20), seed=4), None, 1.0),
    }
    for name in args.ops.split(","):
        opn, opts, gen, cpu_op, alg_factor = cases[name]
        host = gen()
        n = host.nbytes
        d_in = torch.empty(n + 64, dtype=torch.uint8, device=dev)
        d_in[:n].copy_(torch.from_numpy(host))
        torch.cuda.synchronize()
        op = Operator(opn, opts, device=0)
        ext = torch.cuda.ExternalStream(op.stream(), device=dev)
        for _ in range(args.warmup):
            op.reset()
            out = op.call_device(d_in.data_ptr(), n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        launches = 0
        e0.record(ext)
        for _ in range(args.steps):
            op.reset()
            out = op.call_device(d_in.data_ptr(), n)
            launches += op.timings()["kernel_launches"]
        e1.record(ext)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        n_rec, n_out = int(out.n_records), int(out.n)
        alg = n * alg_factor if alg_factor else n + n_out + (16 * n_rec if name == "rmdup" else 0)
        if name == "locate":
            alg = n  # rows are negligible
        line = {"op": name, "operator": opn, "opts": {k: ("1000 x 12-mer" if k == "Pattern" and len(v) > 8 else v) for k, v in opts.items()},
                "in_bytes": n, "out_bytes": n_out, "records": n_rec, "ms_per_step": ms, "records_per_s": n_rec / ms * 1e3,
                "gb_per_s": n / ms / 1e6, "algorithmic_bytes": alg, "whole_step_frac_of_hbm_peak": alg / ms / 1e6 / peak,
                "gpu_launches_per_step": launches // args.steps, "fused_blocks": op.timings()["fused_blocks"]}
        if args.cpu:
            # the oracle port of the reference CPU path on a bounded sample of the same input: all host threads where
            # the port has a sharded driver (seq / stats / rmdup), one thread otherwise
            cap = (256 << 20) if cpu_op else ((4 << 20) if name == 'locate' else (16 << 20))
            sample = np.ascontiguousarray(host[: min(n, cap)])
            k = sample.tobytes().rfind(b"\n@" if sample[0] == 0x40 else b"\n>")
            sample = np.ascontiguousarray(sample[: k + 1])
            if cpu_op:
                threads = os.cpu_count() or 1
                t0 = time.perf_counter()
                nr, _ = oracle.run_mt(cpu_op, sample.ctypes.data, sample.nbytes, opts, threads)
                dt = time.perf_counter() - t0
            else:
                threads = 1
                sb = sample.tobytes()
                dt, _, _ = oracle.time_c_call({"Translate": "orc_translate", "Locate": "orc_locate", "SeqTransform": "orc_seq",
                                               "SubseqTransform": "orc_subseq", "Grep": "orc_grep", "Fq2Fa": "orc_fq2fa"}[opn], sb, opts)
                nr = len(oracle.frame(sb)) - 1
            line["cpu_port"] = {"records_per_s": nr / dt, "gb_per_s": sample.nbytes / dt / 1e9, "cores": threads,
                                "sample_bytes": int(sample.nbytes)}
            line["speedup_vs_cpu_port_gbps"] = (n / ms / 1e6) / (sample.nbytes / dt / 1e9)
        print(json.dumps(line), flush=True)
        op.close()
        del d_in


if __name__ == "__main__":
    main()
