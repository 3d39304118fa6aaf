"""Debug aid: run seq -r -p on a synthetic FASTQ prefix through the C ABI, diff against the oracle, report the first difference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bigseqkit_b200 import Operator, synth
size = int(sys.argv[1]) if len(sys.argv) > 1 else (4 << 20)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
data = synth.fastq_reads(size, seed=2).tobytes()
opts = {"Reverse": True, "Complement": True}
exp, exp_off = oracle.seq(data, opts)
for rep in range(reps):
    with Operator("SeqTransform", opts, device=0) as op:
        got = op.call(data)
        t = op.timings()
    ok_d = got.data == exp
    ok_o = list(got.elem_off) == exp_off
    print("rep", rep, "data_ok", ok_d, "off_ok", ok_o, "n", len(got.data), len(exp), "n_elem", len(got.elem_off) - 1, len(exp_off) - 1,
          "fused", t["fused_blocks"], "launches", t["kernel_launches"], flush=True)
    if not ok_d:
        g = got.data
        m = min(len(g), len(exp))
        i = next((k for k in range(m) if g[k] != exp[k]), m)
        print(" first diff at", i, "tile(out)", i // 16384, "got", g[max(0, i - 40):i + 60], "exp", exp[max(0, i - 40):i + 60])
        nd = sum(1 for k in range(0, m) if g[k] != exp[k])
        print(" differing bytes:", nd)
    if not ok_o:
        go = list(got.elem_off)
        j = next((k for k in range(min(len(go), len(exp_off))) if go[k] != exp_off[k]), -1)
        print(" first off diff at", j, go[j - 2:j + 3] if j >= 0 else None, exp_off[j - 2:j + 3] if j >= 0 else None, len(go), len(exp_off))
