#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (FASTA/FASTQ records/s, GB/s) on the B200 path, next to the reference CPU path.

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (libbsk.so through the C ABI)
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port, all host threads)

Headline (top-level keys) = BASELINE configs[1]: `seq --reverse --complement` on synthetic 150 bp FASTQ, one 1 GiB
block per step per GPU (100 GB = 100 such steps).  `value` times the step with the block resident in HBM; `e2e` times
bsk_run_buffer from pinned host memory (H2D + kernels + D2H).  The `ops` map carries the other configs the metric
names, each measured the same way on its own BASELINE shape:

  stats      C1  stats on 10 M x 100 bp FASTA             stats_all  stats -a on the C2 FASTQ block
  rmdup      C3  rmdup -s, FASTQ with 20 % duplicates     translate  C5  translate -f 6 on CDS-like FASTA
  locate     C4  locate, 1000 x 12-mer panel on contigs

Before anything is timed, the WHOLE block's output of every workload (bytes + element offsets, or the stats row) is
compared with the oracle's output on the same block (sharded over the host threads; that run is also the
`cpu_baseline` of the workload); a mismatch aborts the run.  With --gpus N > 1 every rank processes its own block
(weak scaling); `stats` and `rmdup` then include their exchange step over NCCL inside the timed region
(bsk_stats_allreduce / bsk_rmdup_sharded), checked against the one-rank answer on a prefix of the blocks.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OPTS = {"Reverse": True, "Complement": True}
METRIC = "fastq_records_per_sec"
PROFILES = os.path.join(ROOT, "profiles")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block-mib", type=int, default=1024)
    ap.add_argument("--blocks", type=int, default=3, help="distinct blocks the seq headline cycles through (SURVEY 8d: a unique "
                    "block per step, as consecutive blocks of the 100 GB stream would be)")
    ap.add_argument("--ops", default="stats,stats_all,rmdup,translate,locate",
                    help="extra workloads reported under 'ops' (comma list, 'none' to skip)")
    ap.add_argument("--ops-only", action="store_true", help="profiling aid: skip the seq headline, print only the 'ops' map")
    ap.add_argument("--no-parity", action="store_true", help="profiling aid (ncu runs): skip the oracle comparison")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-block-mib", type=int, default=0, help="pipeline block of bsk_run_buffer (0 = library default)")
    return ap.parse_args()


# ---------------------------------------------------------------------------- workloads
def panel():
    from bigseqkit_b200 import synth
    return synth.pattern_panel(1000, 12, 40)


def workloads(block_mib):
    """name -> (operator, options, generator(buf, rank) -> (array, records), oracle op, config text, dominant kernel)"""
    from bigseqkit_b200 import synth
    blk = block_mib << 20
    c1_rec = 10_000_000 if block_mib >= 1024 else (blk // 114)
    return {
        "seq": ("SeqTransform", OPTS, lambda b, r, j=0: synth.native_fastq(blk, seed=2 + r + 100 * j, out=b), "seq",
                "seq --reverse --complement, synthetic 4-line FASTQ 150 bp (BASELINE configs[1])", "k_fastq_inplace"),
        "stats": ("Stats", {"Tabular": True}, lambda b, r: synth.native_fasta_reads(c1_rec * 114, seed=1 + r, max_records=c1_rec, out=b),
                  "stats", "stats, 10 M x 100 bp single-line FASTA (BASELINE configs[0])", "k_stats_tile"),
        "stats_all": ("Stats", {"Tabular": True, "All": True}, lambda b, r: synth.native_fastq(blk, seed=2 + r, out=b), "stats",
                      "stats -a on the configs[1] FASTQ block", "k_stats_tile"),
        "rmdup": ("RmDup", {"BySeq": True}, lambda b, r: synth.native_fastq(blk, seed=3, dup_frac=0.2, out=b) if r == 0 else
                  _rmdup_block(synth, blk, r, b), "rmdup",
                  "rmdup --by-seq, FASTQ 150 bp with 20 % duplicate sequences (BASELINE configs[2])", "k_rmdup_tile"),
        "translate": ("Translate", {"Frame": ["6"]}, lambda b, r: synth.native_cds(blk, seed=5 + r, out=b), "translate",
                      "translate --frame 6, FASTA CDS 300-3000 bp wrapped at 60 (BASELINE configs[4])", "k_translate"),
        "locate": ("Locate", {"Pattern": panel()}, lambda b, r: synth.native_contigs(min(blk, 256 << 20), seed=4 + r, out=b), "locate",
                   "locate, 1000 x 12-mer panel, FASTA contigs log-uniform 1 kb - 5 Mb wrapped at 60 (BASELINE configs[3]; "
                   "256 MiB block: the reference algorithm the oracle restates makes 2000 passes per contig)", "k_locate_tile"),
    }


def _rmdup_block(synth, blk, rank, buf):
    """rank r > 0: its own reads (10 % in-block copies) of which every 10th record carries a sequence of rank 0's
    block, so that the union over ranks has cross-rank duplicates to find (about 20 % duplicates overall)"""
    import numpy as np
    arr, n = synth.native_fastq(blk, seed=3 + 1000 * rank, dup_frac=0.1, out=buf)
    donor, nd = synth.native_fastq(min(blk, 128 << 20), seed=3, dup_frac=0.2)
    d_nl = np.flatnonzero(donor == 10)
    a_nl = np.flatnonzero(arr == 10)
    # sequence line of record i = bytes (nl[4i] + 1 .. nl[4i + 1])
    k = min(n // 10, nd)
    src = d_nl[0:4 * k:4] + 1
    dst = a_nl[0:40 * k:40] + 1
    idx = np.arange(150)
    arr[(dst[:, None] + idx[None, :]).ravel()] = donor[(src[:, None] + idx[None, :]).ravel()]
    return arr, n


def alg_bytes(name, n, out_bytes, n_rec):
    """algorithmic bytes of the whole operator per block: what the path must move at least"""
    if name == "seq":
        return 2.0 * n
    if name in ("stats", "stats_all", "locate"):
        return float(n)
    if name == "rmdup":
        return float(n + out_bytes + 16 * n_rec)
    return float(n + out_bytes)  # translate: input + proteins


def kernel_alg_bytes(name, n, out_bytes, n_rec):
    """algorithmic bytes of the DOMINANT KERNEL per launch (the kernel `roofline.kernel` names): the same as the
    operator's except for rmdup, whose streaming pass (k_rmdup_tile) reads the block and leaves a 32-byte slot per
    record; the survivors are written by a separate byte-range copy that `whole_step_frac` accounts for"""
    if name == "rmdup":
        return float(n + 32 * n_rec)
    return alg_bytes(name, n, out_bytes, n_rec)


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture
    (profiles/ncu_traffic.json: {workload: {"bytes": ..., "source": ...}}); None when there is no capture"""
    try:
        t = json.load(open(os.path.join(PROFILES, "ncu_traffic.json")))
        return t.get(name)
    except Exception:  # noqa: BLE001
        return None


# ---------------------------------------------------------------------------- helpers
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def pcie_probe(torch, dev, barrier, mib=256, reps=4):
    """pinned-memory DMA rates of this rank with ALL ranks copying at the same time (each leg starts after a barrier):
    H2D alone, D2H alone, both directions at once (GB/s per direction).  The end-to-end path moves every input byte
    H2D and every output byte D2H, so the last figure is its roofline."""
    n = mib << 20
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return n * reps / (time.perf_counter() - t0) / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    return {"h2d_gbs": timed(h2d), "d2h_gbs": timed(d2h), "bidir_gbs_per_direction": timed(both)}


def sha(arr):
    return hashlib.sha256(memoryview(arr)).hexdigest()[:16]


def oracle_full(name, oracle_op, opts, arr, threads):
    """the oracle on the whole block, sharded over the host threads: expected output + its wall time"""
    import oracle
    return oracle.run_mt_full(oracle_op, arr.ctypes.data, arr.nbytes, opts, threads)


def check_parity(name, op, out, exp, is_stats):
    """whole-block comparison of the CUDA output with the oracle's; returns the digest record or raises SystemExit"""
    import numpy as np
    if is_stats:
        row = op.stats_render()
        if row != exp["row"]:
            raise SystemExit("bench.py: %s: CUDA stats row differs from the oracle\n%s\n%s" % (name, row, exp["row"]))
        return {"checked": "stats row of the whole block", "row_sha256": hashlib.sha256(row.encode()).hexdigest()[:16], "match": True}
    data, offs = op.fetch(out)
    ok = data.nbytes == exp["data"].nbytes and np.array_equal(data, exp["data"])
    ok_off = offs is not None and offs.shape == exp["elem_off"].shape and np.array_equal(offs, exp["elem_off"])
    if not (ok and ok_off):
        raise SystemExit("bench.py: %s: CUDA output differs from the oracle (bytes %s, element offsets %s); refusing to "
                         "report a number" % (name, ok, ok_off))
    return {"checked": "whole block: output bytes + element offsets", "bytes": int(data.nbytes), "elements": int(offs.size - 1),
            "sha256": sha(data), "oracle_sha256": sha(exp["data"]), "match": True}


# ---------------------------------------------------------------------------- reference arm
def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    W = workloads(args.block_mib)
    opn, opts, gen, orc, text, _ = W["seq"]
    arr, n_rec = gen(None, 0)
    threads = os.cpu_count() or 1
    addr = arr.ctypes.data
    for _ in range(min(args.warmup, 1)):
        oracle.run_mt(orc, addr, arr.nbytes, opts, threads)
    t0 = time.perf_counter()
    nrec = 0
    for _ in range(args.steps):
        r, _ = oracle.run_mt(orc, addr, arr.nbytes, opts, threads)
        nrec += r
    dt = time.perf_counter() - t0
    val = nrec / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "records/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config(args, text, arr.nbytes, n_rec),
        "gb_per_s": arr.nbytes * args.steps / dt / 1e9,
        "cpu_baseline": {"value": val, "unit": "records/s", "cores": threads, "kind": "port",
                         "sample": "the whole %d MiB block per step (%d records), oracle/bsk_oracle.c orc_run_mt, record-aligned "
                                   "shards, one thread each; the Go/IgnisHPC reference cannot be built in this image"
                                   % (arr.nbytes >> 20, n_rec)},
        "e2e": {"value": val, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def config(args, text, block_bytes, n_rec):
    return {"workload": "%s, %d MiB block per step per GPU (100 GB = %d steps)" % (text, args.block_mib, round(100e9 / block_bytes)),
            "block_bytes": int(block_bytes), "records_per_block": int(n_rec), "read_len": 150,
            "l2_policy": "input block (>= 1 GiB) and output are each far larger than the 126 MB L2; the timed steps cycle through "
                         "%d distinct blocks resident in HBM" % max(args.blocks, 1),
            "parallelism": "shard-per-gpu x%d, no data-path collective" % args.gpus}


# ---------------------------------------------------------------------------- our arm
def bind_numa(local):
    """run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001
        pass
    return None


def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bigseqkit_b200 import Operator
    from bigseqkit_b200 import dist as bd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    numa_cpus = bind_numa(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    threads = os.cpu_count() or 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    W = workloads(args.block_mib)
    names = ([] if args.ops_only else ["seq"]) + [x for x in args.ops.split(",") if x and x != "none"]
    for x in names:
        if x not in W:
            raise SystemExit("bench.py: unknown workload %r" % x)
    cap = max(args.block_mib << 20, 1_140_000_000) + (1 << 20)
    h_in = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_np = h_in.numpy()
    d_buf = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory() if world > 1 else None  # survivors of the sharded rmdup
    cp_stream = torch.cuda.Stream(device=dev)
    if args.e2e_block_mib:
        os.environ["BSK_BLOCK_BYTES"] = str(args.e2e_block_mib << 20)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    sampler = ClockSampler(local)
    sampler.start()
    pcie = None
    if not args.no_e2e:
        p = pcie_probe(torch, dev, barrier)
        agg = allsum([p["h2d_gbs"], p["d2h_gbs"], p["bidir_gbs_per_direction"]])
        mn = [-x for x in allmax([-p["h2d_gbs"], -p["d2h_gbs"], -p["bidir_gbs_per_direction"]])]
        pcie = {"ranks_copying_at_once": world, "h2d_gbs": mn[0], "d2h_gbs": mn[1], "bidir_gbs_per_direction": mn[2],
                "aggregate_h2d_gbs": agg[0], "aggregate_d2h_gbs": agg[1], "aggregate_bidir_gbs_per_direction": agg[2],
                "note": "slowest rank / sum over ranks of a pinned cudaMemcpyAsync probe with all ranks copying together"}

    results = {}
    for name in names:
        opn, opts, gen, orc, text, kernel = W[name]
        is_stats = opn == "Stats"
        collective = world > 1 and name in ("stats", "stats_all", "rmdup")
        arr, n_rec_gen = gen(h_np, rank)
        n = arr.nbytes
        d_buf[:n].copy_(h_in[:n], non_blocking=True)
        torch.cuda.synchronize()
        op = Operator(opn, opts, device=local)
        if collective:
            bd.init_comm(op)
        ext = torch.cuda.ExternalStream(op.stream(), device=dev)

        # the seq headline walks distinct blocks (block j of this rank: another seed), all resident in HBM
        blocks = [(d_buf, n, n_rec_gen)]
        if name == "seq" and args.blocks > 1:
            for j in range(1, args.blocks):
                a_j, nr_j = gen(None, rank, j)
                t_j = torch.empty(a_j.nbytes + 64, dtype=torch.uint8, device=dev)
                t_j[:a_j.nbytes].copy_(torch.from_numpy(a_j))
                if rank == 0 and not args.no_parity:  # every block is checked against the oracle, not only the first
                    exp_j = oracle_full(name, orc, opts, a_j, threads)
                    check_parity(name, op, op.call_device(t_j.data_ptr(), a_j.nbytes), exp_j, False)
                    del exp_j
                blocks.append((t_j, a_j.nbytes, nr_j))
                del a_j
            torch.cuda.synchronize()
        step_no = [0]

        def step():
            if len(blocks) > 1:
                t_b, n_b, _ = blocks[step_no[0] % len(blocks)]
                step_no[0] += 1
                return op.call_device(t_b.data_ptr(), n_b)
            if is_stats:
                op.reset()
                o = op.call_device(d_buf.data_ptr(), n)
                if collective:
                    op.stats_allreduce()
                return o
            if collective:
                return op.rmdup_sharded(d_buf.data_ptr(), n)
            return op.call_device(d_buf.data_ptr(), n)

        # ---- parity gate over the WHOLE block (rank 0; local operator only), and the CPU baseline it doubles as
        parity = cpu = None
        if rank == 0 and not args.no_parity:
            exp = oracle_full(name, orc, opts, arr, threads)
            if is_stats:
                op.reset()
            out = op.call_device(d_buf.data_ptr(), n)
            parity = check_parity(name, op, out, exp, is_stats)
            recs = n_rec_gen
            secs = exp["seconds"]
            how = "the parity run itself (output kept)"
            if name not in ("locate", "translate"):
                # cheap operators: keeping and concatenating the output would dominate; time the plain sharded run
                import oracle
                t0 = time.perf_counter()
                oracle.run_mt(orc, arr.ctypes.data, n, opts, threads)
                secs = time.perf_counter() - t0
                how = "a second sharded run that does not keep the output (orc_run_mt)"
            del exp
            cpu = {"value": recs / secs, "unit": "records/s", "cores": threads, "kind": "port",
                   "gb_per_s": n / secs / 1e9, "seconds": secs,
                   "sample": "the whole block (%d MiB, %d records) once, oracle port of the reference CPU path sharded over "
                             "%d host threads; timed on %s" % (n >> 20, recs, threads, how)}
        # ---- exchange step checked against the one-rank answer on a prefix of every rank's block
        coll_check = None
        if collective:
            coll_check = check_collective(name, opn, opts, orc, arr, op, d_buf, dev, rank, world, local, torch, dist, np)
        barrier()

        for _ in range(max(args.warmup, 3)):
            out = step()
        n_rec = n_rec_gen
        step_no[0] = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_ms = index_ms = op_ms = 0.0
        launches = main_launches = fused = 0
        e0.record(ext)
        for _ in range(args.steps):
            out = step()
            t = op.timings()
            main_ms += t["main_ms"]
            index_ms += t["index_ms"]
            op_ms += t["op_ms"]
            launches += t["kernel_launches"]
            main_launches += t["main_launches"]
            fused += t["fused_blocks"]
        e1.record(ext)
        barrier()
        ms = e0.elapsed_time(e1)
        out_bytes = int(out.n)

        # ---- e2e: pinned host in, pinned host out, through bsk_run_buffer (+ the exchange step at N > 1)
        e2e_s = 0.0
        d2h = 0
        if not args.no_e2e:
            def e2e_step():
                if name == "rmdup" and collective:
                    # host -> HBM, exchange over NCCL, survivors -> host
                    with torch.cuda.stream(cp_stream):  # a torch-owned stream: pinned tensors must not be tied to the ctx stream
                        d_buf[:n].copy_(h_in[:n], non_blocking=True)
                    cp_stream.synchronize()
                    o = op.rmdup_sharded(d_buf.data_ptr(), n)
                    op._check(op.lib.cdll.bsk_memcpy_d2h(op.h, h_out.data_ptr(), o.data, o.n))
                    return int(o.n)
                if is_stats:
                    op.reset()
                r = op.call((h_in.data_ptr(), n), copy=False)
                if is_stats and collective:
                    op.stats_allreduce()
                if is_stats:
                    return len(op.stats_render())
                return int(r.n) + 8 * (int(r.n_elem) + 1)
            for _ in range(2):
                d2h = e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                d2h = e2e_step()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
        ms_max, e2e_max = allmax([ms, e2e_s])
        rec_all, bytes_all = allsum([n_rec, n])
        # records / bytes of the timed steps (the blocks differ by a few records)
        rec_timed = sum(blocks[i % len(blocks)][2] for i in range(args.steps))
        byt_timed = sum(blocks[i % len(blocks)][1] for i in range(args.steps))
        rec_timed_all, byt_timed_all = allsum([rec_timed, byt_timed])
        op.close()

        if rank == 0:
            alg = alg_bytes(name, n, out_bytes, n_rec)
            kalg = kernel_alg_bytes(name, n, out_bytes, n_rec)
            main_per = main_ms / max(main_launches, 1)
            step_ms = ms_max / args.steps
            kern_ms = main_per if main_per > 0 else step_ms
            achieved = kalg / (kern_ms * 1e-3) / 1e9
            traffic = ncu_traffic(name)
            res = {
                "workload": text, "block_bytes": int(n), "records_per_block": int(n_rec), "out_bytes_per_step": out_bytes,
                "value": rec_timed_all / (ms_max * 1e-3), "unit": "records/s",
                "gb_per_s": byt_timed_all / (ms_max * 1e-3) / 1e9, "ms_per_step": step_ms, "distinct_blocks": len(blocks),
                "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic["bytes"] if traffic else None,
                             "traffic_source": traffic["source"] if traffic else None,
                             "algorithmic_bytes_per_launch": kalg, "operator_bytes_per_step": alg, "kernel_ms": kern_ms,
                             "peak_source": peak_src,
                             "whole_step_frac": alg / (step_ms * 1e-3) / 1e9 / peak,
                             "stage_ms": {"index": index_ms / args.steps, "op": op_ms / args.steps}},
                "gpu_launches": int(launches), "fused_blocks": int(fused), "parity": parity,
            }
            if collective:
                res["exchange"] = {"step": "bsk_stats_allreduce (NCCL all-reduce of the dense length histogram + sums)"
                                   if is_stats else "bsk_rmdup_sharded (NCCL all-gather of 16-byte fingerprints, resolve against earlier ranks)",
                                   "inside_timed_region": True, "check": coll_check}
            if not args.no_e2e:
                e2e_gbs = bytes_all * args.steps / e2e_max / 1e9
                res["e2e"] = {"value": rec_all * args.steps / e2e_max, "unit": "records/s", "gb_per_s": e2e_gbs,
                              "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": int(d2h),
                              "frac_of_pcie_probe": e2e_gbs / world / pcie["bidir_gbs_per_direction"] if (pcie and d2h > n // 2)
                              else (e2e_gbs / world / pcie["h2d_gbs"] if pcie else None)}
            if cpu and not args.no_cpu_baseline:
                res["cpu_baseline"] = cpu
            results[name] = res
    sampler.stop_flag = True
    sampler.join()

    if rank == 0 and args.ops_only:
        print(json.dumps({"ops": results, "clocks": sampler.summary()}))
    elif rank == 0:
        s = results["seq"]
        n, n_rec = s["block_bytes"], s["records_per_block"]
        line = {
            "metric": METRIC, "value": s["value"], "unit": "records/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config(args, W["seq"][4], n, n_rec),
            "gb_per_s": s["gb_per_s"], "roofline": s["roofline"], "gpu_launches": s["gpu_launches"],
            "out_bytes_per_step": s["out_bytes_per_step"], "clocks": sampler.summary(), "parity_checked": True,
            "parity": s["parity"],
        }
        line["roofline"]["kernel"] = ("k_fastq_inplace (TMA-staged tile kernel: newline scan + record grammar + in-place revcomp, "
                                      "bulk store)" if s["fused_blocks"] else "k_emit (general path record formatter)")
        if "e2e" in s:
            line["pcie_roofline"] = pcie
            line["e2e"] = dict(s["e2e"], api="bsk_run_buffer (pinned host in -> pinned host out, element offsets included; H2D / "
                               "kernels / D2H pipelined over %s MiB record-aligned blocks on three streams)" % (args.e2e_block_mib or 64))
        if "cpu_baseline" in s:
            line["cpu_baseline"] = s["cpu_baseline"]
        if numa_cpus:
            line["numa"] = {"cpus_bound_per_rank": numa_cpus}
        line["ops"] = {k: v for k, v in results.items() if k != "seq"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def check_collective(name, opn, opts, orc, arr, op, d_buf, dev, rank, world, local, torch, dist, np):
    """N-rank exchange step on a 4 MiB record-aligned prefix of every rank's block against the ONE-rank oracle answer
    on the concatenation of those prefixes (rank order == input order)."""
    import oracle
    cut = 4 << 20
    pre = arr[:cut]
    nl = np.flatnonzero(pre[-4096:] == 10)
    marker = 0x40 if arr[0] == 0x40 else 0x3E
    k = cut
    for j in nl[::-1]:
        p = cut - 4096 + int(j) + 1
        if p < cut and arr[p] == marker and not (marker == 0x40 and arr[p - 2] == 0x2B and arr[p - 3] == 10):
            k = p
            break
    pre = np.ascontiguousarray(arr[:k])
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, torch.tensor([k], dtype=torch.int64, device=dev))
    sizes = [int(x) for x in sizes.tolist()]
    mx = max(sizes)
    mine = torch.zeros(mx, dtype=torch.uint8, device=dev)
    mine[:k] = torch.from_numpy(pre).to(dev)
    allp = torch.zeros(world * mx, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allp, mine)
    d_pre = torch.zeros(k + 64, dtype=torch.uint8, device=dev)
    d_pre[:k] = mine[:k]
    torch.cuda.synchronize()
    if opn == "Stats":
        op.reset()
        op.call_device(d_pre.data_ptr(), k)
        op.stats_allreduce()
        got = op.stats_render()
        ok = True
        if rank == 0:
            allh = allp.cpu().numpy()
            cat = np.concatenate([allh[r * mx:r * mx + sizes[r]] for r in range(world)])
            exp = oracle.run_mt_full("stats", cat.ctypes.data, cat.nbytes, opts, 4)["row"]
            ok = got == exp
            if not ok:
                raise SystemExit("bench.py: %s: %d-rank all-reduced stats row differs from the one-rank oracle row" % (name, world))
        op.reset()
        return {"prefix_bytes_per_rank": int(k), "match": bool(ok)}
    out = op.rmdup_sharded(d_pre.data_ptr(), k)
    kept, _ = op.fetch(out)
    ksz = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(ksz, torch.tensor([kept.nbytes], dtype=torch.int64, device=dev))
    ksz = [int(x) for x in ksz.tolist()]
    kmx = max(max(ksz), 1)
    km = torch.zeros(kmx, dtype=torch.uint8, device=dev)
    km[:kept.nbytes] = torch.from_numpy(kept).to(dev)
    allk = torch.zeros(world * kmx, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allk, km)
    ok = True
    removed = None
    if rank == 0:
        allh, allkh = allp.cpu().numpy(), allk.cpu().numpy()
        cat = np.concatenate([allh[r * mx:r * mx + sizes[r]] for r in range(world)])
        got = np.concatenate([allkh[r * kmx:r * kmx + ksz[r]] for r in range(world)])
        exp = oracle.run_mt_full("rmdup", cat.ctypes.data, cat.nbytes, opts, 4)
        ok = got.nbytes == exp["data"].nbytes and np.array_equal(got, exp["data"])
        removed = int(exp["records"] - (exp["elem_off"].size - 1))
        if not ok:
            raise SystemExit("bench.py: %s: %d-rank union differs from the one-rank oracle answer" % (name, world))
    return {"prefix_bytes_per_rank": int(k), "match": bool(ok), "duplicates_removed_in_prefixes": removed}


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
