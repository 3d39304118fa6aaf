#!/usr/bin/env python
"""Attribute an ncu SASS source page to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [min_pct]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the same cubin by instruction order."""
import csv, io, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ci, cs, csmp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
ins = [(r[cs].strip(), int(r[ci]), int(r[csmp])) for r in rows[hi + 1:] if len(r) > ci and r[ci].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# find the function section
lines = dis.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l]
if not start:
    sys.exit("kernel not found in cubin")
cur, seq = None, []
for l in lines[start[0] + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if seq:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        seq.append(cur)
if len(ins) > len(seq) and len(ins) % len(seq) == 0:
    ins = ins[:len(seq)]  # several launches in the report: attribute the first one
if len(seq) != len(ins):
    print("warning: %d sass rows vs %d disasm instructions" % (len(ins), len(seq)), file=sys.stderr)
agg = {}
tot = sum(i[1] for i in ins)
tots = sum(i[2] for i in ins)
for (txt, n, smp), loc in zip(ins, seq):
    a = agg.setdefault(loc, [0, 0])
    a[0] += n
    a[1] += smp
print("total warp instructions %d, samples %d" % (tot, tots))
src_cache = {}
def src(loc):
    if not loc:
        return ""
    f, ln = loc
    if f not in src_cache:
        try:
            import glob
            p = glob.glob("/root/repo/bigseqkit_b200/csrc/**/" + f, recursive=True)
            src_cache[f] = open(p[0]).read().splitlines() if p else []
        except Exception:
            src_cache[f] = []
    L = src_cache[f]
    return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
# optional buckets: BUCKETS="name:lo-hi,name:lo-hi" over lines of the kernel's main file
import os
if os.environ.get("BUCKETS"):
    bk = []
    for part in os.environ["BUCKETS"].split(","):
        nm, rg = part.split(":")
        lo, hi = rg.split("-")
        bk.append((nm, int(lo), int(hi)))
    bt = {nm: [0, 0] for nm, _, _ in bk}
    bt["other"] = [0, 0]
    for loc, (n, smp) in agg.items():
        nm = "other"
        if loc and loc[0].startswith("k_"):
            for b, lo, hi in bk:
                if lo <= loc[1] <= hi:
                    nm = b
                    break
        bt[nm][0] += n
        bt[nm][1] += smp
    for nm, (n, smp) in bt.items():
        print("BUCKET %-14s %5.1f%% inst  %5.1f%% samples" % (nm, 100.0 * n / tot, 100.0 * smp / max(tots, 1)))
for loc, (n, smp) in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if 100.0 * n / tot >= min_pct or 100.0 * smp / max(tots, 1) >= min_pct:
        print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100.0 * n / tot, 100.0 * smp / max(tots, 1), loc[0] if loc else "?", loc[1] if loc else "?", src(loc)))
