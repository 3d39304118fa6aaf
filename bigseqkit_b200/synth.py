"""Deterministic synthetic FASTA/FASTQ inputs for the BASELINE.json configs.

All generators are seeded numpy (PCG64) and return ``numpy.uint8`` arrays holding whole
records ('\\n'-terminated, no '\\r').  ``nbytes`` is a target: generation stops at the last
whole record that fits.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _digits(vals, width):
    """[R, width] ascii digits of vals (zero padded) and the mask of the non-padding digits."""
    vals = np.asarray(vals, dtype=np.int64)
    pw = 10 ** np.arange(width - 1, -1, -1, dtype=np.int64)
    d = (vals[:, None] // pw[None, :]) % 10
    keep = np.cumsum(d != 0, axis=1) > 0
    keep[:, -1] = True
    return (d + 48).astype(np.uint8), keep


def fastq_reads(nbytes, read_len=150, seed=2, dup_frac=0.0, n_frac=0.001):
    """C2 / C3: 4-line FASTQ, header ``@SIM:1:FC:<lane>:<tile>:<x>:<y> 1:N:0:ACGT``, ``read_len`` bases
    uniform over ACGT with ``n_frac`` N, Phred+33 qualities uniform in [2, 40].  With ``dup_frac`` > 0 each
    record copies, with that probability, the SEQUENCE of a uniformly chosen earlier record (C3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    max_rec = 10 + 1 + 1 + 4 + 1 + 5 + 1 + 5 + 11 + 1 + read_len + 3 + read_len + 1
    R = max(1, int(nbytes // (max_rec - 4)))
    pre = np.frombuffer(b"@SIM:1:FC:", dtype=np.uint8)
    post = np.frombuffer(b" 1:N:0:ACGT\n", dtype=np.uint8)
    lane, lk = _digits(rng.integers(1, 9, R), 1)
    tile, tk = _digits(rng.integers(1101, 2229, R), 4)
    x, xk = _digits(rng.integers(1, 30000, R), 5)
    y, yk = _digits(rng.integers(1, 99999, R), 5)
    colon = np.full((R, 1), ord(":"), np.uint8)
    one = np.ones((R, 1), bool)
    seq = _ACGT[rng.integers(0, 4, (R, read_len), dtype=np.uint8)]
    if n_frac > 0:
        seq[rng.random((R, read_len)) < n_frac] = ord("N")
    if dup_frac > 0 and R > 1:
        is_dup = rng.random(R) < dup_frac
        is_dup[0] = False
        src = (rng.random(R) * np.arange(R)).astype(np.int64)  # uniform in [0, i)
        # resolve chains so that a copy of a copy equals the original
        for i in np.nonzero(is_dup)[0]:
            seq[i] = seq[src[i]]
    qual = rng.integers(2 + 33, 41 + 33, (R, read_len), dtype=np.uint8)
    plus = np.frombuffer(b"\n+\n", dtype=np.uint8)
    nl = np.full((R, 1), 10, np.uint8)
    parts = [np.broadcast_to(pre, (R, pre.size)), lane, colon, tile, colon, x, colon, y,
             np.broadcast_to(post, (R, post.size)), seq, np.broadcast_to(plus, (R, 3)), qual, nl]
    keeps = [np.ones((R, pre.size), bool), lk, one, tk, one, xk, one, yk, np.ones((R, post.size), bool),
             np.ones((R, read_len), bool), np.ones((R, 3), bool), np.ones((R, read_len), bool), one]
    arr = np.concatenate(parts, axis=1)
    keep = np.concatenate(keeps, axis=1)
    out = arr[keep]
    if out.size > nbytes:
        # cut at the last whole record
        rec_len = keep.sum(axis=1)
        ends = np.cumsum(rec_len)
        k = int(np.searchsorted(ends, nbytes, side="right"))
        out = out[: ends[k - 1]] if k > 0 else out[: ends[0]]
    return np.ascontiguousarray(out)


def fasta_reads(n_records, read_len=100, seed=1, width=0):
    """C1: ``>seq%08d`` + ``read_len`` uniform ACGT bases, single line (width=0) or wrapped at ``width``."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R = int(n_records)
    ids, _ = _digits(np.arange(R), 8)
    pre = np.frombuffer(b">seq", dtype=np.uint8)
    nl = np.full((R, 1), 10, np.uint8)
    seq = _ACGT[rng.integers(0, 4, (R, read_len), dtype=np.uint8)]
    if width and width < read_len:
        cols = []
        for s in range(0, read_len, width):
            cols.append(seq[:, s:s + width])
            cols.append(nl)
        body = np.concatenate(cols, axis=1)
    else:
        body = np.concatenate([seq, nl], axis=1)
    arr = np.concatenate([np.broadcast_to(pre, (R, 4)), ids, nl, body], axis=1)
    return np.ascontiguousarray(arr.reshape(-1))


def _wrap_block(seq, width):
    """wrap a 1-D base array at ``width`` columns, trailing newline included."""
    L = seq.size
    nl_count = (L + width - 1) // width
    out = np.empty(L + nl_count, np.uint8)
    idx = np.arange(L)
    out[idx + idx // width] = seq
    ends = np.minimum((np.arange(nl_count) + 1) * width, L) + np.arange(nl_count)
    out[ends] = 10
    return out


def fasta_contigs(nbytes, seed=4, min_len=1000, max_len=5_000_000, width=60):
    """C4: contigs with log-uniform lengths in [min_len, max_len], wrapped at ``width``, uniform ACGT."""
    rng = np.random.Generator(np.random.PCG64(seed))
    chunks, total, i = [], 0, 0
    while True:
        L = int(np.exp(rng.uniform(np.log(min_len), np.log(max_len))))
        need = L + L // width + 32
        if total + need > nbytes and chunks:
            break
        if total + need > nbytes:
            L = max(min_len, int((nbytes - total - 32) * width // (width + 1)))
        hdr = np.frombuffer((">contig%06d len=%d\n" % (i, L)).encode(), dtype=np.uint8)
        seq = _ACGT[rng.integers(0, 4, L, dtype=np.uint8)]
        body = _wrap_block(seq, width)
        chunks += [hdr, body]
        total += hdr.size + body.size
        i += 1
        if total >= nbytes:
            break
    return np.ascontiguousarray(np.concatenate(chunks))


def pattern_panel(n=1000, k=12, seed=40):
    """C4 panel: ``n`` distinct random ``k``-mers over ACGT."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seen, out = set(), []
    while len(out) < n:
        p = bytes(_ACGT[rng.integers(0, 4, k)]).decode()
        if p not in seen:
            seen.add(p)
            out.append(p)
    return out


def fasta_cds(nbytes, seed=5, min_len=300, max_len=3000, width=60):
    """C5: CDS-like records, lengths uniform multiples of 3 in [min_len, max_len], start with ATG, wrapped."""
    rng = np.random.Generator(np.random.PCG64(seed))
    avg = (min_len + max_len) / 2
    R = max(1, int(nbytes / (avg * (1 + 1 / width) + 16)))
    lens = rng.integers(min_len // 3, max_len // 3 + 1, R) * 3
    chunks, total = [], 0
    bases = _ACGT[rng.integers(0, 4, int(lens.sum()), dtype=np.uint8)]
    pos = 0
    for i in range(R):
        L = int(lens[i])
        seq = bases[pos:pos + L].copy()
        pos += L
        seq[:3] = np.frombuffer(b"ATG", dtype=np.uint8)
        hdr = np.frombuffer((">cds%07d gene=g%d\n" % (i, i)).encode(), dtype=np.uint8)
        body = _wrap_block(seq, width)
        if total + hdr.size + body.size > nbytes and chunks:
            break
        chunks += [hdr, body]
        total += hdr.size + body.size
    return np.ascontiguousarray(np.concatenate(chunks))


# ---------------------------------------------------------------------------- native generators (libbsksynth.so)
# Counter-based C generators of the same four shapes (synth_native.c): well under a second per GiB, independent of
# the thread count, able to write straight into a caller buffer (e.g. pinned memory).  The benches and the
# full-size tests use these; the numpy generators above stay for the small parity cases.
_native = None


def _native_lib():
    global _native
    if _native is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbsksynth.so")
        if not os.path.exists(path):
            raise ImportError("bigseqkit_b200: %s is missing; run `make -C bigseqkit_b200/csrc`" % path)
        L = C.CDLL(path)
        u64, u32, vp = C.c_uint64, C.c_uint32, C.c_void_p
        for f, at in (("bsk_synth_fastq", [vp, u64, u64, u32, u32, C.c_int, C.POINTER(u64)]),
                      ("bsk_synth_fasta_reads", [vp, u64, u64, u32, u64, C.c_int, C.POINTER(u64)]),
                      ("bsk_synth_contigs", [vp, u64, u64, u32, u32, u32, C.c_int, C.POINTER(u64)]),
                      ("bsk_synth_cds", [vp, u64, u64, u32, u32, u32, C.c_int, C.POINTER(u64)])):
            getattr(L, f).argtypes = at
            getattr(L, f).restype = u64
        _native = L
    return _native


def _native_call(fn, nbytes, out, args, threads):
    import ctypes as C
    import os
    if out is None:
        out = np.empty(int(nbytes), np.uint8)
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"] and out.nbytes >= nbytes
    nrec = C.c_uint64(0)
    threads = threads or min(os.cpu_count() or 1, 32)
    n = getattr(_native_lib(), fn)(out.ctypes.data, int(nbytes), *args, threads, C.byref(nrec))
    return out[:n], int(nrec.value)


def native_fastq(nbytes, seed=2, read_len=150, dup_frac=0.0, out=None, threads=None):
    """C2 / C3 FASTQ reads (see synth_native.c); returns (uint8 array of whole records <= nbytes, record count)."""
    return _native_call("bsk_synth_fastq", nbytes, out, (int(seed), int(read_len), int(round(dup_frac * 1e6))), threads)


def native_fasta_reads(nbytes, seed=1, read_len=100, max_records=0, out=None, threads=None):
    """C1 FASTA reads, one line per sequence."""
    return _native_call("bsk_synth_fasta_reads", nbytes, out, (int(seed), int(read_len), int(max_records)), threads)


def native_contigs(nbytes, seed=4, min_len=1000, max_len=5_000_000, width=60, out=None, threads=None):
    """C4 FASTA contigs, log-uniform lengths, wrapped."""
    return _native_call("bsk_synth_contigs", nbytes, out, (int(seed), int(min_len), int(max_len), int(width)), threads)


def native_cds(nbytes, seed=5, min_len=300, max_len=3000, width=60, out=None, threads=None):
    """C5 CDS-like FASTA records, lengths multiples of 3, wrapped."""
    return _native_call("bsk_synth_cds", nbytes, out, (int(seed), int(min_len), int(max_len), int(width)), threads)
