#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (FASTA/FASTQ records/s, GB/s) on configs[1]:
`seq --reverse --complement` on synthetic 150 bp FASTQ, one block per step per GPU.

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (libbsk.so through the C ABI)
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port, all host threads)

A "step" is one pass of the hot path (delimiter scan -> record index -> revcomp -> format) over one
block of synthetic input (default 1 GiB, ~3.1 M reads; 100 GB = 100 such steps).  `value` times the
step with the block already resident in HBM; `e2e` times bsk_run_buffer from pinned host memory
(H2D + kernels + D2H of the records and their element offsets).  Shards are independent, so N GPUs
process N different blocks with no data-path collective (weak scaling).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OPTS = {"Reverse": True, "Complement": True}
# measured DRAM traffic of k_fastq_inplace per launch by block size in MiB (ncu, see profiles/); algorithmic = 2 * block
NCU_DRAM_TRAFFIC = {1024: 1078597120 + 1037827584}
METRIC = "fastq_records_per_sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block-mib", type=int, default=1024)
    ap.add_argument("--cpu-sample-mib", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-block-mib", type=int, default=0, help="pipeline block of bsk_run_buffer (0 = library default, 64 MiB)")
    return ap.parse_args()


def config(args, block_bytes, n_rec):
    return {"workload": "seq --reverse --complement, synthetic 4-line FASTQ 150 bp (BASELINE configs[1]), "
                        "%d MiB block per step per GPU (100 GB = %d steps)" % (args.block_mib, round(100e9 / block_bytes)),
            "block_bytes": int(block_bytes), "records_per_block": int(n_rec), "read_len": 150,
            "l2_policy": "input block (>= 1 GiB) and output are each far larger than the 126 MB L2",
            "parallelism": "shard-per-gpu x%d, no data-path collective" % args.gpus}


def cpu_port(sample, threads, steps=1, warmup=0):
    """the oracle port of the reference CPU path (parse -> revcomp -> format, record at a time), multi-threaded"""
    import oracle
    addr = sample.ctypes.data
    for _ in range(warmup):
        oracle.run_mt("seq", addr, sample.nbytes, OPTS, threads)
    t0 = time.perf_counter()
    nrec = 0
    for _ in range(steps):
        r, _ = oracle.run_mt("seq", addr, sample.nbytes, OPTS, threads)
        nrec += r
    dt = time.perf_counter() - t0
    return nrec, dt


def aligned_prefix(arr, nbytes):
    """longest prefix of whole records not exceeding nbytes (synthetic headers start with '@SIM:')"""
    if nbytes >= arr.nbytes:
        return arr
    lo = max(0, nbytes - 4096)
    k = arr[lo:nbytes].tobytes().rfind(b"\n@SIM:")
    return arr[: lo + k + 1]


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


def pcie_probe(torch, dev, mib=256, reps=4):
    """pinned-memory DMA rates on this box: H2D alone, D2H alone, and both directions at once (GB/s per direction).
    The end-to-end path moves every input byte H2D and every output byte D2H, so the last figure is its roofline."""
    n = mib << 20
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
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


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np  # noqa: F401
    from bigseqkit_b200 import synth
    block_bytes = args.block_mib << 20
    sample_bytes = min(args.cpu_sample_mib << 20, block_bytes)
    sample = synth.fastq_reads(sample_bytes, seed=2)
    threads = os.cpu_count() or 1
    nrec, dt = cpu_port(sample, threads, steps=args.steps, warmup=min(args.warmup, 1))
    val = nrec / dt
    n_per = nrec // max(args.steps, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "records/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config(args, block_bytes, n_per * block_bytes // max(sample.nbytes, 1)),
        "gb_per_s": sample.nbytes * args.steps / dt / 1e9,
        "cpu_baseline": {"value": val, "unit": "records/s", "cores": threads, "kind": "port",
                         "sample": "%d MiB prefix of the same synthetic FASTQ per step (%d records), oracle/bsk_oracle.c "
                                   "orc_run_mt, record-aligned shards, one thread each; the Go/IgnisHPC reference cannot "
                                   "be built in this image" % (sample.nbytes >> 20, n_per)},
        "e2e": {"value": val, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bigseqkit_b200 import Operator, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    block_bytes = args.block_mib << 20
    host_np = synth.fastq_reads(block_bytes, seed=2 + rank)
    n = host_np.nbytes
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.numpy()[:] = host_np
    d_in = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    d_in[:n].copy_(h_in, non_blocking=True)
    torch.cuda.synchronize()

    op = Operator("SeqTransform", OPTS, device=local)
    ext = torch.cuda.ExternalStream(op.stream(), device=dev)

    # parity gate on a small prefix (outside the timed region): the oracle is only the checker here
    if rank == 0:
        import oracle
        pre = aligned_prefix(host_np, 4 << 20).tobytes()
        with Operator("SeqTransform", OPTS, device=local) as chk:
            got = chk.call(pre)
        exp, exp_off = oracle.seq(pre, OPTS)
        if got.data != exp or list(got.elem_off) != exp_off:
            raise SystemExit("bench.py: CUDA output differs from the oracle; refusing to report a number")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return op.call_device(d_in.data_ptr(), n)

    for _ in range(max(args.warmup, 3)):
        out = step()
    n_rec = int(out.n_records)
    sampler = ClockSampler(local)
    sampler.start()
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

    # ---- e2e: pinned host in, pinned host out, through bsk_run_buffer
    e2e = None
    pcie = None
    if not args.no_e2e:
        pcie = pcie_probe(torch, dev) if rank == 0 else None
        if args.e2e_block_mib:
            os.environ["BSK_BLOCK_BYTES"] = str(args.e2e_block_mib << 20)
        for _ in range(2):
            r = op.call((h_in.data_ptr(), n), copy=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = op.call((h_in.data_ptr(), n), copy=False)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e = {"s": e2e_s, "d2h": int(r.n) + 8 * (int(r.n_elem) + 1)}
    sampler.stop_flag = True
    sampler.join()

    tt = torch.tensor([ms, e2e["s"] if e2e else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max, e2e_max = float(tt[0]), float(tt[1])
    tot = torch.tensor([n_rec, n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    rec_all, bytes_all = float(tot[0]), float(tot[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        alg_bytes = 2.0 * n  # N read + N written (seq -r -p on FASTQ)
        main_per = main_ms / max(main_launches, 1)
        achieved = alg_bytes / (main_per * 1e-3) / 1e9 if main_per > 0 else 0.0
        step_ms = ms_max / args.steps
        line = {
            "metric": METRIC, "value": rec_all * args.steps / (ms_max * 1e-3), "unit": "records/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config(args, n, n_rec), "gb_per_s": bytes_all * args.steps / (ms_max * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": ("k_fastq_inplace (TMA-staged tile kernel: newline scan + record grammar + in-place revcomp, bulk store)"
                                                    if fused else "k_emit (general path record formatter)"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_TRAFFIC.get(args.block_mib) if fused else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu capture in "
                                           "profiles/r1_ncu_dram_traffic_k_fastq_inplace_1GiB.csv (same block size)",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": main_per, "peak_source": peak_src,
                         "whole_step_frac": alg_bytes / (step_ms * 1e-3) / 1e9 / peak,
                         "stage_ms": {"index": index_ms / args.steps, "op": op_ms / args.steps}},
            "gpu_launches": int(launches), "out_bytes_per_step": out_bytes, "clocks": sampler.summary(),
            "parity_checked": True,
        }
        if e2e:
            e2e_gbs = bytes_all * args.steps / e2e_max / 1e9
            if pcie:
                pcie["frac_of_bidir"] = e2e_gbs / world / pcie["bidir_gbs_per_direction"]
            line["pcie_roofline"] = pcie
            line["e2e"] = {"value": rec_all * args.steps / e2e_max, "unit": "records/s",
                           "gb_per_s": bytes_all * args.steps / e2e_max / 1e9,
                           "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": e2e["d2h"],
                           "api": "bsk_run_buffer (pinned host in -> pinned host out, element offsets included; H2D / kernels / D2H "
                                  "pipelined over %s MiB record-aligned blocks on three streams)" % (args.e2e_block_mib or 64)}
        if not args.no_cpu_baseline and world == 1:
            sample = np.ascontiguousarray(aligned_prefix(host_np, min(args.cpu_sample_mib << 20, n)))
            threads = os.cpu_count() or 1
            nrec_c, dt = cpu_port(sample, threads)
            line["cpu_baseline"] = {"value": nrec_c / dt, "unit": "records/s", "cores": threads, "kind": "port",
                                    "gb_per_s": sample.nbytes / dt / 1e9,
                                    "sample": "%d MiB prefix of the step's block (%d records), oracle port of the reference "
                                              "CPU path, one thread per record-aligned shard" % (sample.nbytes >> 20, nrec_c)}
        print(json.dumps(line))
    op.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
