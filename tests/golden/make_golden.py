#!/usr/bin/env python
"""Regenerates tests/golden/pins.json -- every pin that exists for the hot path.

The reference (citiususc/BigSeqKit @3ab4862) ships no tests or fixtures and cannot be built or
imported offline, so the pins are:
  * the two known-answer tables in the reference's own help text, PARSED HERE FROM THE REFERENCE TREE
    (bigseqkit-cli/helper.go:348-361 `regionExample`; bigseqkit-cli/translate.go:42-52 ambiguous
    codons; translate.go:55-78 the list of genetic-code ids);
  * XXH64 (cespare/xxhash/v2 v2.1.2 == the public XXH64, seed 0) answers produced by the independent
    Python `xxhash` package for lengths that cover every tail branch (32-byte stripes, 8/4/1 tails);
  * the hand-derived record-level vectors of SURVEY.md 4.3 (derived by reading bigseqkit-lib/*.go).

Run in the build container (needs /root/reference and the `xxhash` package):
    python tests/golden/make_golden.py
"""
import json
import os
import random
import re

import xxhash

REF = os.environ.get("BSK_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def region_table():
    src = open(os.path.join(REF, "bigseqkit-cli", "helper.go")).read()
    block = src.split("var regionExample = `", 1)[1].split("`", 1)[0]
    seq, rows = None, {}
    for line in block.splitlines():
        m = re.match(r"\s*seq\s+(.*)$", line)
        if m:
            seq = m.group(1).replace(" ", "")
            continue
        m = re.match(r"\s*(-?\d+:-?\d+)\s+(.*)$", line)
        if m:
            rows[m.group(1)] = m.group(2).replace(" ", "")
    assert seq == "ACGTNacgtn" and len(rows) == 9, (seq, rows)
    return {"source": "bigseqkit-cli/helper.go:348-361", "seq": seq, "regions": rows}


def translate_tables():
    src = open(os.path.join(REF, "bigseqkit-cli", "translate.go")).read()
    amb = dict(re.findall(r"^\s+([ACGTNMRYKSWBDHV]{3}) -> ([A-Z*])\s*$", src, re.M))
    ids = [int(x) for x in re.findall(r"^\s+(\d+): [A-Z]", src, re.M)]
    assert len(amb) == 10 and len(ids) == 24, (amb, ids)
    return {"source": "bigseqkit-cli/translate.go:42-52,55-78", "table": 1, "ambiguous_codons": amb, "table_ids": ids}


def xxh64_vectors():
    rng = random.Random(64)
    vecs = [("", None), ("a", None), ("ACGT", None)]
    for n in list(range(0, 72)) + [95, 96, 97, 127, 128, 150, 151, 255, 256, 1000, 4099]:
        vecs.append(("".join(rng.choice("ACGTNacgtn") for _ in range(n)), None))
    out = []
    for s, _ in vecs:
        h = xxhash.xxh64(s.encode(), seed=0).intdigest()
        out.append({"subject": s, "u64": "%016x" % h, "go_int64": h - (1 << 64) if h >= (1 << 63) else h})
    return {"source": "python xxhash %s, XXH64 seed 0 == cespare/xxhash/v2 Sum64 (bigseqkit-lib/rmdup.go:67-86)" % xxhash.VERSION,
            "vectors": out}


# SURVEY.md 4.3 (hand-derived from bigseqkit-lib/*.go; output stream = element + "\n", helper.go:447)
RECORD_KATS = [
    {"op": "SeqTransform", "opts": {"Reverse": True, "Complement": True}, "in": "@r1 d\nACGTN\n+\nIIIJK\n",
     "out": "@r1 d\nNACGT\n+\nKJIII\n", "cite": "lib/seq.go:188-196"},
    {"op": "SeqTransform", "opts": {"Config": {"LineWidth": 4}}, "in": ">s1\nACGTAC\nGT\n", "out": ">s1\nACGT\nACGT\n",
     "cite": "lib/helper.go:240-250, lib/seq.go:244"},
    {"op": "SeqTransform", "opts": {"RemoveGaps": True}, "in": ">s1\nAC-G T\n", "out": ">s1\nACGT\n", "cite": "bigseqkit/seq.go:41"},
    {"op": "SeqTransform", "opts": {"Name": True}, "in": "@r\nAC\n+\nII\n", "out": "r\n", "cite": "lib/seq.go:151-163"},
    {"op": "SeqTransform", "opts": {"Seq": True}, "in": "@r\nAC\n+\nII\n", "out": "AC\n", "cite": "lib/seq.go:181-184"},
    {"op": "SeqTransform", "opts": {"Qual": True}, "in": "@r\nAC\n+\nII\n", "out": "II\n", "cite": "lib/seq.go:252-259"},
    {"op": "RmDup", "opts": {"BySeq": True}, "in": "@a\nACGT\n+\nIIII\n@b\nACGT\n+\nJJJJ\n@c\nAGGT\n+\nIIII\n",
     "out": "@a\nACGT\n+\nIIII\n@c\nAGGT\n+\nIIII\n", "cite": "lib/rmdup.go:180-215"},
    {"op": "Locate", "opts": {"Pattern": ["TAC", "GTA"]}, "in": ">c\nGATTACA\n",
     "out": "seqID\tpatternName\tpattern\tstrand\tstart\tend\tmatched\nc\tTAC\tTAC\t+\t4\t6\tTAC\nc\tGTA\tGTA\t-\t4\t6\tGTA\n",
     "cite": "lib/locate.go:583-766"},
    {"op": "Locate", "opts": {"Pattern": ["AA"], "OnlyPositiveStrand": True}, "in": ">c\nAAAA\n",
     "out": "seqID\tpatternName\tpattern\tstrand\tstart\tend\tmatched\nc\tAA\tAA\t+\t1\t2\tAA\nc\tAA\tAA\t+\t2\t3\tAA\nc\tAA\tAA\t+\t3\t4\tAA\n",
     "cite": "lib/locate.go:659-663"},
    {"op": "Locate", "opts": {"Pattern": ["AA"], "OnlyPositiveStrand": True, "NonGreedy": True}, "in": ">c\nAAAA\n",
     "out": "seqID\tpatternName\tpattern\tstrand\tstart\tend\tmatched\nc\tAA\tAA\t+\t1\t2\tAA\n", "cite": "lib/locate.go:660"},
    {"op": "Translate", "opts": {"Frame": ["6"]}, "in": ">x\nATGGCCTAA\n",
     "out": ">x\nMA*\n>x\nWP\n>x\nGL\n>x\nLGH\n>x\n*A\n>x\nRP\n", "cite": "lib/translate.go:106-142"},
    {"op": "SubseqTransform", "opts": {"Region": "2:-2"}, "in": ">x\nACGTNacgtn\n", "out": ">x\nCGTNacgt\n",
     "cite": "lib/subseq.go:189-190,314-317"},
    {"op": "Grep", "opts": {"Pattern": ["b"]}, "in": ">a\nAC\n>b\nGG\n>c\nTT\n", "out": ">b\nGG\n", "cite": "lib/grep.go:501-512"},
    {"op": "Grep", "opts": {"Pattern": ["GTA"], "BySeq": True}, "in": ">x\nGATTACA\n", "out": ">x\nGATTACA\n",
     "cite": "lib/grep.go:442-482"},
]
STATS_KATS = [
    {"opts": {"Tabular": True}, "in": ">a\nACGT\n>b\nACGTAC\n>c\nGGGTTT\n",
     "out": "file\tformat\ttype\tnum_seqs\tsum_len\tmin_len\tavg_len\tmax_len\ninput0\tN/A\tDNA\t3\t16\t4\t5.3\t6\n",
     "cite": "bigseqkit/stats.go:199-207, bigseqkit-cli/stats.go:17"},
]

# NCBI genetic codes (public constants, base order TCAG) for the ids the reference lists
NCBI_TABLE_1 = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"

if __name__ == "__main__":
    pins = {"region": region_table(), "translate": translate_tables(), "xxh64": xxh64_vectors(),
            "records": RECORD_KATS, "stats": STATS_KATS, "ncbi_table_1": NCBI_TABLE_1}
    with open(os.path.join(HERE, "pins.json"), "w") as f:
        json.dump(pins, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote pins.json: %d xxh64 vectors, %d record vectors" % (len(pins["xxh64"]["vectors"]), len(RECORD_KATS)))
