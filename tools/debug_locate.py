#!/usr/bin/env python
"""Diff of the locate tile path against the oracle on the first test input of tests/test_locate_tile.py (rows missing / extra)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from bigseqkit_b200.api import Operator
import test_locate_tile as T

rng = random.Random(1060)
data = T._fasta(rng, [30000, 5, 0, 41000, 12, 977], 60, n_frac=0.002)
pats = T._panel(rng, data, 12, 40)
exp, _ = oracle.locate(data, {"Pattern": pats})
with Operator("Locate", {"Pattern": pats}) as op:
    got = op.call(data)
    print("fused blocks", op.timings()["fused_blocks"])
e, g = set(exp.split(b"\n")), set(bytes(got.data).split(b"\n"))
print("expected rows", len(e), "got", len(g), "missing", len(e - g), "extra", len(g - e))
for r in sorted(e - g)[:8]: print("  missing", r)
for r in sorted(g - e)[:8]: print("  extra  ", r)
