"""A second, independent restatement (pure Python, small inputs only) of the parts of the hot path that can be written
in a few lines: framing + SeqParser.Read, seq --reverse --complement, the stats row, rmdup --by-seq (keys from the
Python xxhash package) and translate with the standard table.  It exists to cross-check the C oracle: two
restatements written separately from the same reference lines must agree."""
import math

import xxhash

DNA_PAIRS = dict(zip(b"acgtryswkmbdhvnACGTRYSWKMBDHVN", b"tgcayrswmkvhdbnTGCAYRSWMKVHDBN"))
NCBI1 = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "U": "T", "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC",
         "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}


def frame(data):
    """record start offsets: FASTA lines starting '>'; FASTQ lines starting '@' unless the previous line is a bare '+'"""
    if not data:
        return []
    fq = data[:1] == b"@"
    marker = b"@" if fq else b">"
    starts = [0]
    pos = 0
    while True:
        i = data.find(b"\n" + marker, pos)
        if i < 0:
            break
        if not (fq and i >= 2 and data[i - 2:i] == b"\n+"):
            starts.append(i + 1)
        pos = i + 1
    return starts


def records(data):
    """(head, seq, qual, is_fastq) per record, SeqParser.Read (lib/helper.go:219-325) with the marker stripped"""
    st = frame(data) + [len(data)]
    fq = data[:1] == b"@"
    out = []
    for a, b in zip(st[:-1], st[1:]):
        e = data[a:b]
        if e.endswith(b"\n"):
            e = e[:-1]
        if e[:1] in (b"@", b">"):
            e = e[1:]
        lines = e.split(b"\n")
        head = lines[0]
        if not fq:
            out.append((head, b"".join(lines[1:]), b"", False))
            continue
        seq, qual, in_qual = [], [], False
        terminated = lines[1:-1] if len(lines) > 1 else []
        last = lines[-1] if len(lines) > 1 else None
        for ln in terminated:
            if not in_qual and ln[:1] == b"+":
                in_qual = True
            elif in_qual:
                qual.append(ln)
            else:
                seq.append(ln)
        if last is not None and in_qual:  # the unterminated final segment only counts in quality mode
            qual.append(last)
        out.append((head, b"".join(seq), b"".join(qual), True))
    return out


def seq_revcomp(data):
    out = []
    for head, s, q, fq in records(data):
        rc = bytes(DNA_PAIRS.get(c, c) for c in reversed(s))
        if fq:
            out.append(b"@" + head + b"\n" + rc + b"\n+\n" + q[::-1] + b"\n")
        else:
            wrapped = b"\n".join(rc[i:i + 60] for i in range(0, len(rc), 60))
            out.append(b">" + head + b"\n" + wrapped + b"\n")
    return b"".join(out)


def stats_row(data):
    lens = [len(s) for _, s, _, _ in records(data)]
    n, tot = len(lens), sum(lens)
    avg = math.floor(tot / n * 10 + 0.5) / 10 if n else 0.0
    return "input0\tN/A\tDNA\t%d\t%d\t%d\t%.1f\t%d" % (n, tot, min(lens) if n else 0, avg, max(lens) if n else 0)


def rmdup_by_seq(data):
    seen, out, keys = set(), [], []
    for head, s, q, fq in records(data):
        h = xxhash.xxh64(s, seed=0).intdigest()
        keys.append(h - (1 << 64) if h >= (1 << 63) else h)
        if s in seen:
            continue
        seen.add(s)
        out.append((b"@" + head + b"\n" + s + b"\n+\n" + q + b"\n") if fq else
                   (b">" + head + b"\n" + b"\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + b"\n"))
    return b"".join(out), keys


def translate_frame1(data):
    out = []
    for head, s, _, _ in records(data):
        prot = []
        u = s.decode().upper()
        for i in range(0, len(u) - 2, 3):
            opts = set()
            for a in IUPAC.get(u[i], "?"):
                for b in IUPAC.get(u[i + 1], "?"):
                    for c in IUPAC.get(u[i + 2], "?"):
                        opts.add(NCBI1["TCAG".index(a) * 16 + "TCAG".index(b) * 4 + "TCAG".index(c)] if "?" not in a + b + c else "X")
            prot.append(opts.pop() if len(opts) == 1 else "X")
        p = "".join(prot).encode()
        out.append(b">" + head + b"\n" + b"\n".join(p[i:i + 60] for i in range(0, len(p), 60)) + b"\n")
    return b"".join(out)
