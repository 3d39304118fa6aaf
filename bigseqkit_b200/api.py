"""ctypes binding of libbsk.so (include/bsk.h) and a thin mirror of the reference's
Python driver surface (bigseqkit-py/bigseqkit/seq.py ... : ``seq(input, o=None, **kwargs)``
where the kwargs are the lower-camel spellings of the Go option fields).

There is no CPU implementation behind this module: without the CUDA library and
a GPU every call raises.
"""
import ctypes as C
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsk.so")

BSK_OK = 0
BSK_ERR_ARG, BSK_ERR_DATA, BSK_ERR_CUDA, BSK_ERR_UNSUPPORTED, BSK_ERR_STATE, BSK_ERR_NOMEM = -1, -2, -3, -4, -5, -6
BSK_COMM_ID_BYTES = 128


class BskError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class _Out(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("elem_off", C.c_void_p), ("n_elem", C.c_size_t),
                ("n_records", C.c_uint64)]


class _Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("num", "sum_len", "min_len", "max_len", "sum_gap", "q20", "q30", "n50", "l50")] + \
               [(k, C.c_double) for k in ("avg_len", "q1", "q2", "q3", "q20_pct", "q30_pct")] + \
               [("type", C.c_char * 16), ("hist_len", C.POINTER(C.c_uint64)), ("hist_cnt", C.POINTER(C.c_uint64)),
                ("n_hist", C.c_size_t)]


class _Timings(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("index_ms", "op_ms", "main_ms", "total_ms")] + \
               [("main_launches", C.c_uint64), ("kernel_launches", C.c_uint64), ("in_bytes", C.c_uint64),
                ("out_bytes", C.c_uint64), ("fused_blocks", C.c_uint64)]


# every symbol include/bsk.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "bsk_version", "bsk_device_count", "bsk_create", "bsk_create_error", "bsk_destroy", "bsk_last_error",
    "bsk_set_elem_offsets", "bsk_set_union", "bsk_reset", "bsk_run_buffer", "bsk_run_device", "bsk_stream", "bsk_get_timings",
    "bsk_stats_result", "bsk_stats_merge", "bsk_stats_add", "bsk_stats_dense_device", "bsk_stats_render",
    "bsk_shard_bounds", "bsk_run_file", "bsk_rmdup_keys", "bsk_rmdup_removed", "bsk_rmdup_dup_seqs", "bsk_rmdup_dup_num", "bsk_rmdup_prepare_device", "bsk_rmdup_resolve_device", "bsk_grep_count",
    "bsk_comm_unique_id", "bsk_comm_error", "bsk_comm_init", "bsk_comm_rank", "bsk_comm_free", "bsk_output_offsets",
    "bsk_stats_allreduce", "bsk_rmdup_sharded", "bsk_reduce", "bsk_rmdup_union", "bsk_memcpy_d2h", "bsk_stage_device",
]


class Library:
    """A loaded libbsk.so with typed prototypes."""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise ImportError("bigseqkit_b200: %s is missing; build it with `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % path)
        self.path = path
        L = self.cdll = C.CDLL(path)
        vp, sz, i64, u64 = C.c_void_p, C.c_size_t, C.c_int64, C.c_uint64
        L.bsk_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
        L.bsk_create_error.restype = C.c_char_p
        L.bsk_destroy.argtypes = [vp]
        L.bsk_destroy.restype = None
        L.bsk_last_error.argtypes = [vp]
        L.bsk_last_error.restype = C.c_char_p
        L.bsk_set_elem_offsets.argtypes = [vp, C.c_int]
        L.bsk_set_union.argtypes = [vp, C.c_int]
        L.bsk_reset.argtypes = [vp]
        L.bsk_run_buffer.argtypes = [vp, vp, sz, i64, C.POINTER(_Out)]
        L.bsk_run_device.argtypes = [vp, vp, sz, i64, C.POINTER(_Out)]
        L.bsk_shard_bounds.argtypes = [C.c_char_p, C.c_int, C.POINTER(u64)]
        L.bsk_run_file.argtypes = [vp, C.c_char_p, u64, u64, i64, C.c_char_p, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.bsk_stream.argtypes = [vp]
        L.bsk_stream.restype = vp
        L.bsk_get_timings.argtypes = [vp, C.POINTER(_Timings)]
        L.bsk_stats_result.argtypes = [vp, C.POINTER(_Stats)]
        L.bsk_stats_merge.argtypes = [vp, vp]
        L.bsk_stats_add.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), sz, u64, u64, u64, C.c_char_p]
        L.bsk_stats_dense_device.argtypes = [vp, vp, sz, C.POINTER(u64)]
        L.bsk_stats_render.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, sz]
        L.bsk_stats_render.restype = C.c_long
        L.bsk_rmdup_keys.argtypes = [vp, C.POINTER(C.POINTER(i64)), C.POINTER(sz)]
        L.bsk_rmdup_removed.argtypes = [vp]
        for f in ("bsk_rmdup_dup_seqs", "bsk_rmdup_dup_num"):
            getattr(L, f).argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(sz)]
        L.bsk_rmdup_removed.restype = u64
        L.bsk_rmdup_prepare_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(u64)]
        L.bsk_rmdup_resolve_device.argtypes = [vp, vp, u64, C.POINTER(_Out)]
        L.bsk_grep_count.argtypes = [vp]
        L.bsk_grep_count.restype = u64
        L.bsk_comm_unique_id.argtypes = [C.c_char_p]
        L.bsk_comm_error.restype = C.c_char_p
        L.bsk_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
        L.bsk_comm_rank.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.bsk_comm_free.argtypes = [vp]
        L.bsk_output_offsets.argtypes = [vp, u64, C.POINTER(u64), C.POINTER(u64)]
        L.bsk_stats_allreduce.argtypes = [vp]
        L.bsk_rmdup_sharded.argtypes = [vp, vp, sz, C.POINTER(_Out)]
        L.bsk_reduce.argtypes = [C.POINTER(vp), C.c_int]
        L.bsk_rmdup_union.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(vp), C.POINTER(sz), C.POINTER(_Out)]
        L.bsk_memcpy_d2h.argtypes = [vp, vp, vp, sz]
        L.bsk_stage_device.argtypes = [vp, vp, sz, C.POINTER(vp)]

    def device_count(self):
        return self.cdll.bsk_device_count()


_default = None


def default_library():
    global _default
    if _default is None:
        _default = Library()
    return _default


def shard_bounds(path, n_shards, lib=None):
    """record-aligned byte ranges of a FASTA/FASTQ file (bsk_shard_bounds): n_shards + 1 offsets"""
    lib = lib or default_library()
    b = (C.c_uint64 * (n_shards + 1))()
    rc = lib.cdll.bsk_shard_bounds(os.fsencode(path), n_shards, b)
    if rc != BSK_OK:
        raise BskError(rc, "bsk_shard_bounds(%s) failed" % path)
    return list(b)


def _as_json(opts):
    if opts is None:
        return b"{}"
    if isinstance(opts, bytes):
        return opts
    if isinstance(opts, str):
        return opts.encode()
    return json.dumps(opts).encode()


class Result:
    """Output of one Call(): ``data`` (every element followed by '\\n'), element offsets, input record count."""

    def __init__(self, data, elem_off, n_records):
        self.data = data
        self.elem_off = elem_off
        self.n_records = n_records

    def elements(self):
        o = self.elem_off
        return [self.data[o[i]:o[i + 1] - 1] for i in range(len(o) - 1)]


class Operator:
    """One reference operator between Before() and After() (e.g. SeqTransform, bigseqkit-lib/seq.go:21-26)."""

    def __init__(self, op, opts=None, device=-1, lib=None):
        self.lib = lib or default_library()
        self.op = op
        h = C.c_void_p()
        rc = self.lib.cdll.bsk_create(op.encode(), _as_json(opts), device, C.byref(h))
        if rc != BSK_OK:
            raise BskError(rc, self.lib.cdll.bsk_create_error().decode(errors="replace"))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.cdll.bsk_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != BSK_OK:
            raise BskError(rc, self.lib.cdll.bsk_last_error(self.h).decode(errors="replace"))

    def set_elem_offsets(self, want):
        self._check(self.lib.cdll.bsk_set_elem_offsets(self.h, 1 if want else 0))

    def set_union(self, on=True):
        """the partitions of the following calls are one dataframe (rmdup dedups across them, range keeps counting)"""
        self._check(self.lib.cdll.bsk_set_union(self.h, 1 if on else 0))

    def reset(self):
        self._check(self.lib.cdll.bsk_reset(self.h))

    def call(self, data, partition_id=0, copy=True):
        """Call() on a partition held in host memory (bytes / bytearray / object with the buffer protocol)."""
        if isinstance(data, bytes):
            ptr, n, keep = C.cast(C.c_char_p(data), C.c_void_p), len(data), data
        elif isinstance(data, tuple):  # (address, nbytes) of caller-managed (e.g. pinned) memory
            ptr, n, keep = C.c_void_p(data[0]), data[1], None
        else:
            mv = memoryview(data).cast("B")
            n = mv.nbytes
            keep = (C.c_uint8 * n).from_buffer(mv) if not mv.readonly else (C.c_uint8 * n).from_buffer_copy(mv)
            ptr = C.cast(keep, C.c_void_p)
        out = _Out()
        self._check(self.lib.cdll.bsk_run_buffer(self.h, ptr, n, partition_id, C.byref(out)))
        del keep
        if not copy:
            return out
        d = C.string_at(out.data, out.n) if out.n else b""
        offs = None
        if out.elem_off:
            offs = list((C.c_uint64 * (out.n_elem + 1)).from_address(out.elem_off))
        return Result(d, offs, out.n_records)

    def call_file(self, path, off=0, length=0, out_path=None, out_off=0, partition_id=0):
        """Call() on the byte range [off, off+length) of a file (length 0 = to the end); the result is written to
        out_path at out_off.  Returns (output bytes, input records, elements)."""
        ob, nr, ne = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.cdll.bsk_run_file(self.h, os.fsencode(path), off, length, partition_id,
                                               os.fsencode(out_path) if out_path else None, out_off, C.byref(ob), C.byref(nr),
                                               C.byref(ne)))
        return ob.value, nr.value, ne.value

    def call_device(self, dev_ptr, nbytes, partition_id=0):
        """Call() on a partition resident in HBM; returns the raw struct with DEVICE pointers."""
        out = _Out()
        self._check(self.lib.cdll.bsk_run_device(self.h, C.c_void_p(dev_ptr), nbytes, partition_id, C.byref(out)))
        return out

    def stream(self):
        return self.lib.cdll.bsk_stream(self.h)

    def timings(self):
        t = _Timings()
        self._check(self.lib.cdll.bsk_get_timings(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in _Timings._fields_}

    # ---- stats
    def stats_result(self):
        s = _Stats()
        self._check(self.lib.cdll.bsk_stats_result(self.h, C.byref(s)))
        d = {k: getattr(s, k) for k in ("num", "sum_len", "min_len", "max_len", "sum_gap", "q20", "q30", "n50", "avg_len",
                                        "q1", "q2", "q3", "q20_pct", "q30_pct")}
        d["type"] = s.type.decode()
        d["hist"] = [(s.hist_len[i], s.hist_cnt[i]) for i in range(s.n_hist)]
        return d

    def stats_render(self, file="input0", fmt="N/A"):
        n = self.lib.cdll.bsk_stats_render(self.h, file.encode(), fmt.encode(), None, 0)
        if n < 0:
            self._check(BSK_ERR_STATE)
        buf = C.create_string_buffer(n + 1)
        self.lib.cdll.bsk_stats_render(self.h, file.encode(), fmt.encode(), buf, n + 1)
        return buf.value.decode()

    def stats_merge(self, other):
        self._check(self.lib.cdll.bsk_stats_merge(self.h, other.h))

    def stats_add(self, hist, q20=0, q30=0, sum_gap=0, type=""):
        n = len(hist)
        a = (C.c_uint64 * max(n, 1))(*[h[0] for h in hist])
        b = (C.c_uint64 * max(n, 1))(*[h[1] for h in hist])
        self._check(self.lib.cdll.bsk_stats_add(self.h, a, b, n, q20, q30, sum_gap, type.encode()))

    def stats_dense_device(self, dev_ptr, nbins):
        over = C.c_uint64(0)
        self._check(self.lib.cdll.bsk_stats_dense_device(self.h, C.c_void_p(dev_ptr), nbins, C.byref(over)))
        return over.value

    # ---- rmdup
    def rmdup_keys(self):
        kp = C.POINTER(C.c_int64)()
        n = C.c_size_t(0)
        self._check(self.lib.cdll.bsk_rmdup_keys(self.h, C.byref(kp), C.byref(n)))
        return [kp[i] for i in range(n.value)]

    def rmdup_removed(self):
        return self.lib.cdll.bsk_rmdup_removed(self.h)

    def _rmdup_text(self, fn):
        p = C.c_void_p()
        n = C.c_size_t(0)
        self._check(getattr(self.lib.cdll, fn)(self.h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value) if n.value else b""

    def rmdup_dup_seqs(self):
        """-d: removed records (bigseqkit-lib/rmdup.go:185-187), accumulated since the partition started."""
        return self._rmdup_text("bsk_rmdup_dup_seqs")

    def rmdup_dup_num(self):
        """-D: "count\tid1, id2, ..." rows (bigseqkit-lib/rmdup.go:230-234)."""
        return self._rmdup_text("bsk_rmdup_dup_num")

    def rmdup_prepare_device(self, dev_ptr, nbytes, fp_ptr, fp_cap):
        nrec = C.c_uint64(0)
        self._check(self.lib.cdll.bsk_rmdup_prepare_device(self.h, C.c_void_p(dev_ptr), nbytes, C.c_void_p(fp_ptr), fp_cap,
                                                           C.byref(nrec)))
        return nrec.value

    def rmdup_resolve_device(self, all_fp_ptr, n_before):
        out = _Out()
        self._check(self.lib.cdll.bsk_rmdup_resolve_device(self.h, C.c_void_p(all_fp_ptr), n_before, C.byref(out)))
        return out

    def grep_count(self):
        return self.lib.cdll.bsk_grep_count(self.h)

    # ---- exchange steps (NCCL communicator bound to the ctx)
    def comm_init(self, uid, n_ranks, rank):
        """uid: the 128 bytes rank 0 got from `comm_unique_id()` and the driver shipped to every rank."""
        assert len(uid) == BSK_COMM_ID_BYTES
        self._check(self.lib.cdll.bsk_comm_init(self.h, bytes(uid), n_ranks, rank))

    def comm_free(self):
        self._check(self.lib.cdll.bsk_comm_free(self.h))

    def comm_rank(self):
        r, n = C.c_int(0), C.c_int(1)
        self.lib.cdll.bsk_comm_rank(self.h, C.byref(r), C.byref(n))
        return r.value, n.value

    def output_offsets(self, n_local):
        off, tot = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.cdll.bsk_output_offsets(self.h, n_local, C.byref(off), C.byref(tot)))
        return off.value, tot.value

    def stats_allreduce(self):
        """StatsReduce over all ranks (bigseqkit/stats.go:91): afterwards this ctx holds the global totals."""
        self._check(self.lib.cdll.bsk_stats_allreduce(self.h))

    def rmdup_sharded(self, dev_ptr, nbytes):
        """rmdup over the shards of all ranks (bigseqkit/rmdup.go:97); returns the raw struct with DEVICE pointers."""
        out = _Out()
        self._check(self.lib.cdll.bsk_rmdup_sharded(self.h, C.c_void_p(dev_ptr), nbytes, C.byref(out)))
        return out

    def fetch(self, out):
        """copy a device-side result (call_device / rmdup_sharded / ...) to the host: Result with bytes + offsets"""
        import numpy as np
        data = np.empty(out.n, np.uint8)
        self._check(self.lib.cdll.bsk_memcpy_d2h(self.h, data.ctypes.data, out.data, out.n))
        offs = None
        if out.elem_off:
            offs = np.empty(out.n_elem + 1, np.uint64)
            self._check(self.lib.cdll.bsk_memcpy_d2h(self.h, offs.ctypes.data, out.elem_off, offs.nbytes))
        return data, offs


def comm_unique_id(lib=None):
    """rank 0: a fresh NCCL unique id (128 bytes) to ship to every rank before Operator.comm_init"""
    lib = lib or default_library()
    buf = C.create_string_buffer(BSK_COMM_ID_BYTES)
    rc = lib.cdll.bsk_comm_unique_id(buf)
    if rc != BSK_OK:
        raise BskError(rc, lib.cdll.bsk_comm_error().decode(errors="replace"))
    return buf.raw


def reduce_local(ops):
    """StatsReduce folded over several Stats operators of this process: every one ends with the sum (bsk_reduce)"""
    lib = ops[0].lib
    arr = (C.c_void_p * len(ops))(*[o.h for o in ops])
    rc = lib.cdll.bsk_reduce(arr, len(ops))
    if rc != BSK_OK:
        raise BskError(rc, lib.cdll.bsk_last_error(ops[0].h).decode(errors="replace"))


def rmdup_union_local(ops, dev_ptrs, sizes):
    """rmdup over several shards held by ctxs of this process, ctx order == input order (bsk_rmdup_union)"""
    lib = ops[0].lib
    n = len(ops)
    arr = (C.c_void_p * n)(*[o.h for o in ops])
    ptrs = (C.c_void_p * n)(*dev_ptrs)
    szs = (C.c_size_t * n)(*sizes)
    outs = (_Out * n)()
    rc = lib.cdll.bsk_rmdup_union(arr, n, ptrs, szs, outs)
    if rc != BSK_OK:
        raise BskError(rc, lib.cdll.bsk_last_error(ops[0].h).decode(errors="replace"))
    return list(outs)


# ---------------------------------------------------------------------------
# driver-style helpers: bigseqkit-py spells options as lower-camel kwargs
_CONFIG_KEYS = {"SeqType", "LineWidth", "IDRegexp", "IDNCBI", "Quiet", "AlphabetGuessSeqLength", "ChunkSize", "BufferSize"}


def make_opts(opts=None, **kwargs):
    """Build the reference JSON dict from a dict and/or kwargs (``reverse=True`` -> ``{"Reverse": true}``;
    KitConfig fields such as ``lineWidth`` go under ``"Config"``)."""
    d = dict(opts or {})
    cfg = dict(d.get("Config") or {})
    for k, v in kwargs.items():
        key = k[0].upper() + k[1:]
        if key in ("IdRegexp",):
            key = "IDRegexp"
        if key in ("IdNcbi", "IdNCBI"):
            key = "IDNCBI"
        if key in _CONFIG_KEYS:
            cfg[key] = v
        else:
            d[key] = v
    if cfg:
        d["Config"] = cfg
    return d


def _run(op, data, opts, kwargs, device=-1, lib=None):
    with Operator(op, make_opts(opts, **kwargs), device=device, lib=lib) as o:
        return o.call(data)


def seq(data, o=None, **kwargs):          # bigseqkit-py seq.py / bigseqkit/seq.go:157-170
    return _run("SeqTransform", data, o, kwargs)


def subseq(data, o=None, **kwargs):       # bigseqkit/subseq.go
    return _run("SubseqTransform", data, o, kwargs)


def fq2fa(data, o=None, **kwargs):        # bigseqkit/fq2fa.go:25-39
    return _run("Fq2Fa", data, o, kwargs)


def translate(data, o=None, **kwargs):    # bigseqkit/translate.go
    return _run("Translate", data, o, kwargs)


def locate(data, o=None, **kwargs):       # bigseqkit/locate.go:122-134
    return _run("Locate", data, o, kwargs)


def grep(data, o=None, **kwargs):         # bigseqkit/grep.go:121-181
    return _run("Grep", data, o, kwargs)


def rmDup(data, o=None, **kwargs):        # bigseqkit/rmdup.go:70-108
    return _run("RmDup", data, o, kwargs)


def stats(data, o=None, file="input0", fmt="N/A", **kwargs):   # bigseqkit/stats.go:75-288 (Stats + StatsString)
    with Operator("Stats", make_opts(o, **kwargs)) as op:
        op.call(data)
        return op.stats_result(), op.stats_render(file, fmt)
