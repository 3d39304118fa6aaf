#!/usr/bin/env python
"""gpurun_out/<tag>_* -> profiles/<tag>_*: bench lines, launch lists, ncu --set full summaries with per-source-line
instruction shares, and profiles/ncu_traffic.json (DRAM read + write bytes of the dominant kernel of every workload,
which bench.py reports as roofline.traffic).  usage: python tools/collect_profiles.py <tag>"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KERNEL = {"seq": ("k_fastq_inplace", "k_fastq_inplace"), "stats": ("k_stats_tile", "k_stats_tileILb0"),
          "stats_all": ("k_stats_tile", "k_stats_tileILb1"), "rmdup": ("k_rmdup_tile", "k_rmdup_tileE"),
          "translate": ("k_translate_tile", "k_translate_tileE"), "locate": ("k_locate_tile", "k_locate_tileE")}
for f in ("bench.json", "bench_ref.json", "file.jsonl", "pytest.log", "smi.txt"):
    src = os.path.join(G, "%s_%s" % (tag, f))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "%s_%s" % (tag, f)))
traffic = {}
for wl, (obj, sym) in KERNEL.items():
    lc = os.path.join(G, "%s_%s_launches.csv" % (tag, wl))
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(P, "%s_ncu_launches_%s.csv" % (tag, wl)))
    rep = os.path.join(G, "%s_%s_prof.ncu-rep" % (tag, wl))
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    d = dict(zip(rows[0], rows[2]))
    u = dict(zip(rows[0], rows[1]))

    def val(k):
        x = float(d[k].replace(",", ""))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u[k], 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    traffic[wl] = {"bytes": rd + wr, "read": rd, "write": wr, "kernel": d.get("Kernel Name", "")[:80],
                   "block": "256 MiB" if wl == "locate" else "1 GiB (stats: 1.14 GB)",
                   "source": "profiles/%s_ncu_%s_summary.txt (ncu --set full, one launch)" % (tag, wl)}
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, obj], capture_output=True, text=True).stdout
    cub = os.path.join(ROOT, "bigseqkit_b200", "csrc", obj + ".sm_100a.cubin")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join("build", obj + ".o")], cwd=os.path.join(ROOT, "bigseqkit_b200", "csrc"),
                   capture_output=True)
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, cub, sym, "0.8"], capture_output=True,
                           text=True, cwd=os.path.join(ROOT, "bigseqkit_b200", "csrc")).stdout
    with open(os.path.join(P, "%s_ncu_%s_summary.txt" % (tag, wl)), "w") as fh:
        fh.write("# ncu --set full --clock-control none, workload '%s' of bench.py, B200 (tools/gpu_round.sh %s)\n" % (wl, tag))
        fh.write(summ)
        fh.write("\n# instruction shares per source line (tools/ncu_lines.py; >= 0.8 %)\n")
        fh.write(lines)
    if os.path.exists(cub):
        os.remove(cub)
if traffic:
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1)[:1500])
