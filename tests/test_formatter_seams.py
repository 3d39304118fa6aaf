"""The output formatters (k_emit, k_emit_contig) against the oracle on inputs shaped to hit their seams: a CTA formats
16 KiB of output from a slice of the record table staged in shared memory (<= 1024 records, views of <= 256), finds the
record of a 16-byte chunk through a chunk -> record map, writes chunks inside one run of source bytes directly and
defers the chunks that cross a piece / record boundary.  Tiny records overflow the slice (global-memory walk), runs of
dropped records leave entries without output in the map, long records span several CTAs."""
import random

import pytest

from util import check_parity


def _fq(rng, lens, name=lambda i: "r%d" % i):
    out = []
    for i, L in enumerate(lens):
        s = "".join(rng.choice("ACGT") for _ in range(L))
        q = "".join(chr(33 + rng.randrange(41)) for _ in range(L))
        out.append("@%s\n%s\n+\n%s\n" % (name(i), s, q))
    return "".join(out).encode()


def _fa(rng, lens, width=0, name=lambda i: "s%d" % i):
    out = []
    for i, L in enumerate(lens):
        s = "".join(rng.choice("ACGT") for _ in range(L))
        body = s if width <= 0 else "\n".join(s[j:j + width] for j in range(0, L, width))
        out.append(">%s\n%s\n" % (name(i), body))
    return "".join(out).encode()


SHAPES = {
    # > 1024 records per 16 KiB of output: the slice does not fit, every chunk walks the global arrays
    "tiny_fasta": lambda rng: _fa(rng, [rng.randrange(1, 4) for _ in range(9000)]),
    "tiny_fastq": lambda rng: _fq(rng, [rng.randrange(1, 3) for _ in range(6000)]),
    # 256 < records <= 1024 per CTA: offsets staged, views from global memory
    "short_fastq": lambda rng: _fq(rng, [rng.randrange(8, 14) for _ in range(3000)]),
    # reads: everything staged
    "reads": lambda rng: _fq(rng, [rng.randrange(140, 152) for _ in range(400)]),
    # records longer than a CTA's 16 KiB, next to tiny ones
    "mixed": lambda rng: _fa(rng, [40000, 1, 2, 17000, 3, 16384, 16, 15, 33000]),
    "mixed_wrapped": lambda rng: _fa(rng, [40000, 1, 2, 17000, 3, 16384, 16, 15, 33000], width=60),
    # names of every length around the 16-byte chunk
    "names": lambda rng: _fq(rng, [rng.randrange(20, 40) for _ in range(300)], name=lambda i: "n" * (i % 37) + str(i)),
}


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("opts", [{}, {"Reverse": True, "Complement": True}, {"MinLen": 2}, {"Name": True}, {"Seq": True}],
                         ids=["plain", "revcomp", "minlen", "name", "seq"])
def test_seq_formatter(lib, shape, opts):
    data = SHAPES[shape](random.Random(sum(shape.encode())))
    check_parity(lib, "SeqTransform", data, opts)


@pytest.mark.parametrize("shape", ["tiny_fasta", "tiny_fastq", "short_fastq", "reads", "mixed"])
def test_rmdup_compaction(lib, shape):
    rng = random.Random(7)
    data = SHAPES[shape](rng)
    # duplicate a third of the records in long runs so that whole CTAs hold only dropped records
    recs = data.split(b"\n@" if data[:1] == b"@" else b"\n>")
    mark = data[:1]
    recs = [recs[0][1:]] + recs[1:]
    recs[-1] = recs[-1].rstrip(b"\n")
    dup = recs[: len(recs) // 3]
    body = recs + dup + recs[len(recs) // 2:]
    data2 = b"".join(mark + r + b"\n" for r in body)
    check_parity(lib, "RmDup", data2, {"BySeq": True})
    check_parity(lib, "RmDup", data2, {})


@pytest.mark.parametrize("shape", ["tiny_fastq", "reads", "names"])
def test_grep_and_subseq(lib, shape):
    data = SHAPES[shape](random.Random(11))
    check_parity(lib, "Grep", data, {"Pattern": ["r7", "r8", "r4000", "n7"], "InvertMatch": True})
    check_parity(lib, "Grep", data, {"Pattern": ["r7", "r8", "r4000", "n7"]})
    if shape != "tiny_fastq":
        check_parity(lib, "SubseqTransform", data, {"Region": "3:-3"})
    check_parity(lib, "Fq2Fa", data, {})
