#!/usr/bin/env python
"""Scope F of SURVEY section 8d: file -> file wall clock of `seq -r -p` through bsk_run_file (reader thread -> pinned
slots -> H2D / kernels / D2H -> writer thread, all overlapped; host memory bounded by six 64 MiB slots), with the file
in the page cache (tmpfs when available).  One JSON line per I/O thread count; the CPU port (oracle, all threads, in
memory) beside it."""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigseqkit_b200 import synth  # noqa: E402
from bigseqkit_b200.api import Operator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--threads", default="1,0")  # 0 = library default (all cores, at most 16)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    data = synth.fastq_reads(a.mib << 20, seed=2)
    src = os.path.join(a.dir, "bsk_bench_in.fq")
    dst = os.path.join(a.dir, "bsk_bench_out.fq")
    data.tofile(src)
    n = data.nbytes
    try:
        for t in a.threads.split(","):
            if int(t) > 0:
                os.environ["BSK_IO_THREADS"] = t
            else:
                os.environ.pop("BSK_IO_THREADS", None)
            with Operator("SeqTransform", {"Reverse": True, "Complement": True}) as op:
                op.set_elem_offsets(False)
                times, nrec, ob = [], 0, 0
                for rep in range(a.reps + 1):  # the first repetition warms the arenas up and is not counted
                    open(dst, "wb").close()
                    t0 = time.perf_counter()
                    ob, nrec, _ = op.call_file(src, 0, 0, dst, 0)
                    if rep:
                        times.append(time.perf_counter() - t0)
                best = min(times)
                tm = op.timings()
            line = {"scope": "file_to_file", "op": "seq -r -p", "io_threads": int(t) or "default", "in_bytes": int(n),
                    "out_bytes": int(ob), "records": int(nrec), "best_s": best, "gb_per_s_in": n / best / 1e9,
                    "records_per_s": nrec / best, "dir": a.dir, "gpu_ms_in_call": tm.get("total_ms")}
            print(json.dumps(line), flush=True)
        if a.cpu:
            import numpy as np
            import oracle
            threads = os.cpu_count() or 1
            opts = {"Reverse": True, "Complement": True}
            sample = synth.fastq_reads(min(n, 256 << 20), seed=2)
            t0 = time.perf_counter()
            nr, _ = oracle.run_mt("seq", sample.ctypes.data, sample.nbytes, opts, threads)
            dt = time.perf_counter() - t0
            print(json.dumps({"scope": "cpu_port_in_memory", "cores": threads, "sample_bytes": int(sample.nbytes),
                              "gb_per_s_in": sample.nbytes / dt / 1e9, "records_per_s": nr / dt}), flush=True)
            # the same port file -> file: read the file, transform on all threads, write the result (what the reference
            # does around its operators: PlainFile ... FileStore)
            best = None
            for rep in range(2):
                open(dst, "wb").close()
                t0 = time.perf_counter()
                arr = np.fromfile(src, dtype=np.uint8)
                res = oracle.run_mt_full("seq", arr.ctypes.data, arr.nbytes, opts, threads)
                res["data"].tofile(dst)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
                nr = res["records"]
                del res, arr
            print(json.dumps({"scope": "cpu_port_file_to_file", "cores": threads, "in_bytes": int(n), "best_s": best,
                              "gb_per_s_in": n / best / 1e9, "records_per_s": nr / best}), flush=True)
    finally:
        for p in (src, dst):
            if os.path.exists(p):
                os.remove(p)


if __name__ == "__main__":
    main()
