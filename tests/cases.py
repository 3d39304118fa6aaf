"""Shared inputs for the parity tests: hand-written edge cases + a seeded fuzz generator."""
import random

FQ_SIMPLE = b"@r1 d\nACGTN\n+\nIIIJK\n@r2\nAAC\n+\nI@I\n"
FA_SIMPLE = b">s1 x\nACGTAC\nGT\n>s2\nAC-G T\n"

EDGE_INPUTS = {
    "empty": b"",
    "fq_simple": FQ_SIMPLE,
    "fa_simple": FA_SIMPLE,
    "fq_no_final_newline": b"@a\nACGT\n+\nIIII\n@b\nGG\n+\nJJ",
    "fq_qual_starts_with_at": b"@a\nACGT\n+\n@III\n@b\nGGCC\n+\n@@@@\n",
    "fq_plus_with_name": b"@a desc\nACGT\n+a desc\nIIII\n@b\nGGCC\n+b\nJJJJ\n",
    "fq_multiline": b"@a\nACGT\nAC\n+\nIIII\nII\n@b\nGG\n+\nJJ\n",
    "fq_blank_line_end": b"@a\nACGT\n+\nIIII\n\n@b\nGG\n+\nJJ\n\n",
    "fq_one_base": b"@a\nA\n+\nI\n@b\nC\n+\nJ\n",
    "fq_empty_seq": b"@a\n\n+\n\n@b\nAC\n+\nII\n",
    "fq_lower_iupac": b"@a x y\nacgtnRYKM\n+\nIIIIIIIII\n@b\tz\nNNNN\n+\n!!!!\n",
    "fa_single": b">only\nACGT",
    "fa_header_only": b">h1\n>h2\nACGT\n>h3\n",
    "fa_no_marker_first": b"ACGT\n>b\nGGTT\n",
    "fa_leading_newline": b"\n>a\nACGT\n>b\nGG\n",
    "fa_blank_lines": b">a\nACGT\n\nAC\n\n>b\n\nGG\n",
    "fa_wrapped": b">a some desc\n" + b"ACGTACGTAC\n" * 7 + b"ACG\n>b\n" + b"TTTTGGGGCC\n" * 3,
    "fa_protein": b">p1\nMKVLAAGIVGLLLAQ*\n>p2\nMEEPQSDPSV\n",
    "fa_rna": b">r1\nACGUACGU\n>r2\nuuuaaccgg\n",
    "fa_gaps": b">g1\nAC-GT..AC GT\n>g2\n----\n",
    "fa_tabs_in_header": b">id1\tdesc one\nACGT\n>id2  two spaces\nGGCC\n> lead\nAA\n",
    "fa_ncbi": b">gi|110645304|ref|NC_002516.2| Pseudomonas\nACGT\n>plain\nGG\n",
    "fa_long_line": b">long\n" + b"ACGTTGCA" * 700 + b"\n>short\nAC\n",
    "fq_mismatch": b"@a\nACGT\n+\nIII\n@b\nGG\n+\nJJ\n",
}


def fuzz_fasta(rng, n_rec=None, alphabet="ACGT", max_len=200, width=None, final_nl=True):
    n_rec = rng.randint(1, 40) if n_rec is None else n_rec
    out = []
    for i in range(n_rec):
        name = "s%d" % i
        if rng.random() < 0.5:
            name += rng.choice([" ", "\t", "  "]) + "desc%d" % rng.randint(0, 99)
        L = rng.choice([0, 1, 2, 3, rng.randint(0, max_len), rng.randint(0, max_len)])
        seq = "".join(rng.choice(alphabet) for _ in range(L))
        w = width if width is not None else rng.choice([0, 0, 7, 10, 60])
        out.append(">" + name + "\n")
        if w and L:
            out.append("\n".join(seq[j:j + w] for j in range(0, L, w)) + "\n")
        else:
            out.append(seq + "\n")
    s = "".join(out)
    if not final_nl and s.endswith("\n"):
        s = s[:-1]
    return s.encode()


def fuzz_fastq(rng, n_rec=None, alphabet="ACGTN", max_len=200, final_nl=True, fixed_len=None):
    n_rec = rng.randint(1, 40) if n_rec is None else n_rec
    out = []
    for i in range(n_rec):
        name = "r%d" % i
        if rng.random() < 0.5:
            name += " " + "x%d" % rng.randint(0, 9)
        L = fixed_len if fixed_len is not None else rng.choice([1, 2, rng.randint(1, max_len), rng.randint(1, max_len)])
        seq = "".join(rng.choice(alphabet) for _ in range(L))
        qual = "".join(chr(rng.randint(33, 74)) for _ in range(L))
        plus = "+" if rng.random() < 0.8 else "+" + name
        out.append("@%s\n%s\n%s\n%s\n" % (name, seq, plus, qual))
    s = "".join(out)
    if not final_nl:
        s = s[:-1]
    return s.encode()


def fuzz_inputs(seed, count=12):
    rng = random.Random(seed)
    res = []
    for k in range(count):
        kind = k % 6
        if kind == 0:
            res.append(fuzz_fasta(rng))
        elif kind == 1:
            res.append(fuzz_fastq(rng))
        elif kind == 2:
            res.append(fuzz_fasta(rng, alphabet="ACGTacgtNRYKMSWBDHV", final_nl=rng.random() < 0.5))
        elif kind == 3:
            res.append(fuzz_fastq(rng, final_nl=rng.random() < 0.5))
        elif kind == 4:
            res.append(fuzz_fasta(rng, alphabet="ACGT-. ", max_len=80))
        else:
            res.append(fuzz_fasta(rng, n_rec=rng.randint(1, 4), max_len=5000, width=60))
    return res
