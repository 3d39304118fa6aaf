"""ctypes binding of oracle/bsk_oracle.c (TEST INFRASTRUCTURE ONLY).

Options are given as the same dict/JSON the reference ships between driver and
executor (bigseqkit/helper.go:47-66): {"Config": {...}, "Reverse": true, ...}.
"""
import ctypes as C
import json
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbskoracle.so")


def build(force=False):
    src = os.path.join(_HERE, "bsk_oracle.c")
    hdr = os.path.join(_HERE, "bsk_oracle.h")
    if (force or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


class _Opts(C.Structure):
    _fields_ = [
        ("SeqType", C.c_char_p), ("LineWidth", C.c_int), ("IDNCBI", C.c_int), ("AlphabetGuessSeqLength", C.c_int),
        ("Reverse", C.c_int), ("Complement", C.c_int), ("Name", C.c_int), ("Seq", C.c_int), ("Qual", C.c_int),
        ("OnlyId", C.c_int), ("RemoveGaps", C.c_int), ("GapLetters", C.c_char_p),
        ("LowerCase", C.c_int), ("UpperCase", C.c_int), ("Dna2rna", C.c_int), ("Rna2dna", C.c_int),
        ("ValidateSeq", C.c_int), ("ValidateSeqLength", C.c_int), ("MaxLen", C.c_int), ("MinLen", C.c_int),
        ("QualAsciiBase", C.c_int), ("MinQual", C.c_double), ("MaxQual", C.c_double),
        ("Tabular", C.c_int), ("All", C.c_int), ("FqEncoding", C.c_char_p),
        ("ByName", C.c_int), ("BySeq", C.c_int), ("IgnoreCase", C.c_int), ("OnlyPositiveStrand", C.c_int),
        ("TranslTable", C.c_int), ("Frame", C.c_char_p),
        ("Trim", C.c_int), ("Clean", C.c_int), ("AllowUnknownCodon", C.c_int), ("InitCodonAsM", C.c_int),
        ("AppendFrame", C.c_int),
        ("n_patterns", C.c_int), ("pattern_names", C.POINTER(C.c_char_p)), ("patterns", C.POINTER(C.c_char_p)),
        ("NonGreedy", C.c_int), ("Gtf", C.c_int), ("Bed", C.c_int), ("HideMatched", C.c_int), ("Circular", C.c_int),
        ("InvertMatch", C.c_int), ("Count", C.c_int),
        ("Region", C.c_char_p),
    ]


class _Out(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("n", C.c_size_t), ("elem_off", C.POINTER(C.c_uint64)),
                ("n_elem", C.c_size_t), ("err", C.c_char * 512)]


class _Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("num", "sum_len", "min_len", "max_len", "sum_gap", "q20", "q30", "n50", "l50")] + \
               [(k, C.c_double) for k in ("avg_len", "q1", "q2", "q3", "q20_pct", "q30_pct")] + \
               [("type", C.c_char * 16), ("hist_len", C.POINTER(C.c_uint64)), ("hist_cnt", C.POINTER(C.c_uint64)),
                ("n_hist", C.c_size_t), ("err", C.c_char * 512)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_opts_default.argtypes = [C.POINTER(_Opts)]
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
        for f in ("orc_seq", "orc_translate", "orc_locate", "orc_grep", "orc_subseq", "orc_fq2fa"):
            getattr(L, f).argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Opts), C.POINTER(_Out)]
        L.orc_rmdup.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Opts), C.POINTER(_Out), C.POINTER(C.c_uint64)]
        L.orc_rmdup_dups.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Opts), C.POINTER(_Out), C.POINTER(_Out)]
        L.orc_rmdup_keys.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Opts), C.POINTER(C.POINTER(C.c_int64)),
                                     C.POINTER(C.c_size_t)]
        L.orc_stats_run.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Opts), C.POINTER(_Stats)]
        L.orc_stats_merge.argtypes = [C.POINTER(_Stats), C.POINTER(_Stats)]
        L.orc_stats_finalise.argtypes = [C.POINTER(_Stats), C.c_int]
        L.orc_stats_render.restype = C.c_void_p
        L.orc_stats_render.argtypes = [C.POINTER(_Stats), C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.orc_frame.restype = C.c_size_t
        L.orc_frame.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint64))]
        L.orc_subseq_range.restype = C.c_size_t
        L.orc_subseq_range.argtypes = [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_translate_codon.argtypes = [C.c_int, C.c_char_p]
        L.orc_wrap_len.restype = C.c_size_t
        L.orc_wrap_len.argtypes = [C.c_size_t, C.c_int]
        L.orc_run_mt.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(_Opts), C.c_int,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_run_mt_out.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(_Opts), C.c_int, C.POINTER(_Out),
                                     C.POINTER(_Stats), C.POINTER(C.c_uint64)]
        L.orc_out_free.argtypes = [C.POINTER(_Out)]
        L.orc_stats_free.argtypes = [C.POINTER(_Stats)]
        L.free = C.CDLL(None).free
        L.free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _b(v):
    return v if isinstance(v, bytes) else str(v).encode()


def _mk_opts(opts, keep):
    """opts: dict in the reference JSON schema (or JSON string).  Pattern lists may be
    [pattern,...] (name = pattern) or [(name, pattern), ...]."""
    if isinstance(opts, (str, bytes)):
        opts = json.loads(opts)
    opts = dict(opts or {})
    o = _Opts()
    lib().orc_opts_default(C.byref(o))
    cfg = opts.pop("Config", None) or {}
    flat = dict(cfg)
    flat.update(opts)
    for k, v in flat.items():
        if v is None:
            continue
        if k in ("Pattern", "PatternFile", "IDRegexp", "Quiet", "ChunkSize", "BufferSize", "SkipErr", "Basename",
                 "DupSeqsFile", "DupNumFile", "UseRegexp", "Degenerate", "UseFmi", "MaxMismatch", "DeleteMatched",
                 "ListTranslTable", "ListTranslTableWithAmbCodons", "Chr", "Gtf_", "Feature", "UpStream", "DownStream",
                 "OnlyFlank", "GtfTag"):
            continue
        if k == "Frame":
            v = ",".join(str(x) for x in v) if isinstance(v, (list, tuple)) else str(v)
        if k == "ValidateSeqLength" and k in cfg and "ValidateSeqLength" in opts:
            v = opts["ValidateSeqLength"]
        if not hasattr(o, k):
            raise KeyError("oracle: unknown option %r" % k)
        cur = getattr(o, k)
        if isinstance(v, (str, bytes)):
            bv = _b(v)
            keep.append(bv)
            setattr(o, k, bv)
        elif isinstance(cur, float) or isinstance(v, float) and k in ("MinQual", "MaxQual"):
            setattr(o, k, float(v))
        else:
            setattr(o, k, int(v))
    pats = flat.get("Pattern") or []
    pats = [p for p in pats if p != "" or len(pats) > 1]
    if pats:
        names, seqs = [], []
        for p in pats:
            if isinstance(p, (tuple, list)):
                names.append(_b(p[0])); seqs.append(_b(p[1]))
            else:
                names.append(_b(p)); seqs.append(_b(p))
        an = (C.c_char_p * len(names))(*names)
        ap = (C.c_char_p * len(seqs))(*seqs)
        keep.extend([names, seqs, an, ap])
        o.n_patterns = len(names)
        o.pattern_names = an
        o.patterns = ap
    return o


def _take(out, rc):
    L = lib()
    data = C.string_at(out.data, out.n) if out.n else b""
    offs = [out.elem_off[i] for i in range(out.n_elem + 1)] if out.elem_off else [0]
    err = out.err.decode(errors="replace")
    L.orc_out_free(C.byref(out))
    if rc != 0:
        raise OracleError(err)
    return data, offs


def _run(fn, data, opts):
    keep = []
    o = _mk_opts(opts, keep)
    out = _Out()
    rc = getattr(lib(), fn)(data, len(data), C.byref(o), C.byref(out))
    return _take(out, rc)


def time_c_call(fn, data, opts):
    """Wall time of the C function alone (no Python conversion of the result); returns (seconds, out bytes, elements).
    Used by the CPU-baseline legs of the benches."""
    import time
    keep = []
    o = _mk_opts(opts, keep)
    out = _Out()
    t0 = time.perf_counter()
    rc = getattr(lib(), fn)(data, len(data), C.byref(o), C.byref(out))
    dt = time.perf_counter() - t0
    n, ne = out.n, out.n_elem
    err = out.err.decode(errors="replace")
    lib().orc_out_free(C.byref(out))
    if rc != 0:
        raise OracleError(err)
    return dt, n, ne


def seq(data, opts=None):
    return _run("orc_seq", data, opts)


def translate(data, opts=None):
    return _run("orc_translate", data, opts)


def locate(data, opts=None):
    return _run("orc_locate", data, opts)


def grep(data, opts=None):
    return _run("orc_grep", data, opts)


def subseq(data, opts=None):
    return _run("orc_subseq", data, opts)


def fq2fa(data, opts=None):
    return _run("orc_fq2fa", data, opts)


def duplicate(data, times):
    out = _Out()
    rc = lib().orc_duplicate(data, len(data), C.c_int64(times), C.byref(out))
    return _take(out, rc)


def range_(data, start, end, index_base=0):
    """RangePrepare + RangeFilter with the 0-based half-open bounds the library operator receives"""
    out = _Out()
    rc = lib().orc_range(data, len(data), C.c_int64(start), C.c_int64(end), C.c_int64(index_base), C.byref(out))
    return _take(out, rc)


def head(data, n):
    return range_(data, 0, n)


def rmdup(data, opts=None):
    keep = []
    o = _mk_opts(opts, keep)
    out = _Out()
    removed = C.c_uint64(0)
    rc = lib().orc_rmdup(data, len(data), C.byref(o), C.byref(out), C.byref(removed))
    d, offs = _take(out, rc)
    return d, offs, removed.value


def rmdup_dups(data, opts=None):
    """(-d text, -D text): removed records and the "count\tid1, id2" rows (lib/rmdup.go:180-239)."""
    keep = []
    o = _mk_opts(opts, keep)
    o1, o2 = _Out(), _Out()
    rc = lib().orc_rmdup_dups(data, len(data), C.byref(o), C.byref(o1), C.byref(o2))
    err = o1.err.decode(errors="replace")
    d2 = C.string_at(o2.data, o2.n) if o2.n else b""
    lib().orc_out_free(C.byref(o2))
    d1, _ = _take(o1, rc)
    return d1, d2


def rmdup_keys(data, opts=None):
    keep = []
    o = _mk_opts(opts, keep)
    kp = C.POINTER(C.c_int64)()
    n = C.c_size_t(0)
    rc = lib().orc_rmdup_keys(data, len(data), C.byref(o), C.byref(kp), C.byref(n))
    keys = [kp[i] for i in range(n.value)]
    lib().free(kp)
    if rc != 0:
        raise OracleError("rmdup_keys failed")
    return keys


def _stats_dict(s):
    d = {k: getattr(s, k) for k in ("num", "sum_len", "min_len", "max_len", "sum_gap", "q20", "q30", "n50",
                                    "avg_len", "q1", "q2", "q3", "q20_pct", "q30_pct")}
    d["type"] = s.type.decode()
    d["hist"] = [(s.hist_len[i], s.hist_cnt[i]) for i in range(s.n_hist)]
    return d


def stats(data, opts=None, file="input0", fmt="N/A"):
    """Returns (dict, rendered StatsString)."""
    keep = []
    o = _mk_opts(opts, keep)
    s = _Stats()
    L = lib()
    rc = L.orc_stats_run(data, len(data), C.byref(o), C.byref(s))
    if rc != 0:
        err = s.err.decode(errors="replace")
        L.orc_stats_free(C.byref(s))
        raise OracleError(err)
    d = _stats_dict(s)
    p = L.orc_stats_render(C.byref(s), _b(file), _b(fmt), o.Tabular, o.All)
    txt = C.string_at(p).decode()
    L.free(p)
    L.orc_stats_free(C.byref(s))
    return d, txt


def stats_sharded(shards, opts=None, file="input0", fmt="N/A"):
    """stats over several shards merged with sum semantics."""
    keep = []
    o = _mk_opts(opts, keep)
    L = lib()
    acc = None
    for sh in shards:
        s = _Stats()
        if L.orc_stats_run(sh, len(sh), C.byref(o), C.byref(s)) != 0:
            raise OracleError(s.err.decode(errors="replace"))
        if acc is None:
            acc = s
        else:
            L.orc_stats_merge(C.byref(acc), C.byref(s))
            L.orc_stats_free(C.byref(s))
    L.orc_stats_finalise(C.byref(acc), o.All)
    d = _stats_dict(acc)
    p = L.orc_stats_render(C.byref(acc), _b(file), _b(fmt), o.Tabular, o.All)
    txt = C.string_at(p).decode()
    L.free(p)
    L.orc_stats_free(C.byref(acc))
    return d, txt


def frame(data):
    sp = C.POINTER(C.c_uint64)()
    n = lib().orc_frame(data, len(data), C.byref(sp))
    st = [sp[i] for i in range(n + 1)]
    lib().free(sp)
    return st


def xxh64(b, seed=0):
    return lib().orc_xxh64(b, len(b), seed)


def subseq_range(length, start, end):
    s0 = C.c_size_t(0)
    n = lib().orc_subseq_range(length, start, end, C.byref(s0))
    return s0.value, n


def translate_codon(table, codon):
    r = lib().orc_translate_codon(table, _b(codon))
    return None if r < 0 else chr(r)


def wrap_len(l, w):
    return lib().orc_wrap_len(l, w)


def run_mt(op, data_ptr, n, opts, threads):
    """data_ptr: integer address (or bytes).  Returns (records, out_bytes)."""
    keep = []
    o = _mk_opts(opts, keep)
    nr, ob = C.c_uint64(0), C.c_uint64(0)
    if isinstance(data_ptr, (bytes, bytearray)):
        buf = (C.c_char * len(data_ptr)).from_buffer_copy(data_ptr)
        keep.append(buf)
        data_ptr = C.addressof(buf)
    rc = lib().orc_run_mt(_b(op), C.c_void_p(data_ptr), n, C.byref(o), threads, C.byref(nr), C.byref(ob))
    if rc != 0:
        raise OracleError("run_mt(%s) failed" % op)
    return nr.value, ob.value


def run_mt_full(op, data_ptr, n, opts, threads, file="input0", fmt="N/A"):
    """Sharded run over the WHOLE input with the output kept (full-size parity checks).  data_ptr: integer address of
    n bytes.  Returns a dict: seconds (the C call alone), records, and for op "stats" the rendered row (`row`), else
    `data` (numpy uint8 view copy of every element + '\n') and `elem_off` (numpy uint64, n_elem + 1)."""
    import time
    import numpy as np
    keep = []
    o = _mk_opts(opts, keep)
    out, st, nr = _Out(), _Stats(), C.c_uint64(0)
    L = lib()
    t0 = time.perf_counter()
    rc = L.orc_run_mt_out(_b(op), C.c_void_p(data_ptr), n, C.byref(o), threads, C.byref(out), C.byref(st), C.byref(nr))
    dt = time.perf_counter() - t0
    if rc != 0:
        err = out.err.decode(errors="replace")
        L.orc_out_free(C.byref(out))
        L.orc_stats_free(C.byref(st))
        raise OracleError("run_mt_full(%s): %s" % (op, err))
    res = {"seconds": dt, "records": nr.value}
    if op == "stats":
        p = L.orc_stats_render(C.byref(st), _b(file), _b(fmt), o.Tabular, o.All)
        res["row"] = C.string_at(p).decode()
        res["stats"] = _stats_dict(st)
        L.free(p)
        L.orc_stats_free(C.byref(st))
    else:
        res["data"] = np.ctypeslib.as_array(out.data, shape=(out.n,)).copy() if out.n else np.zeros(0, np.uint8)
        res["elem_off"] = np.ctypeslib.as_array(out.elem_off, shape=(out.n_elem + 1,)).copy() if out.elem_off else np.zeros(1, np.uint64)
        L.orc_out_free(C.byref(out))
    return res


def elements(data, offs):
    """split an output stream into its elements (each followed by one '\\n')."""
    return [data[offs[i]:offs[i + 1] - 1] for i in range(len(offs) - 1)]
