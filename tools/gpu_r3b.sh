#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== default"; timeout 300 python tools/debug_locate.py 2>&1 | tail -12
echo "== plain"; BSK_LT_PLAIN=1 timeout 300 python tools/debug_locate.py 2>&1 | tail -12
for pl in 0 1; do
BSK_LT_PLAIN=$pl timeout 600 python -m pytest tests/test_locate_tile.py -m gpu -x -q 2>&1 | tail -2
( BSK_LT_PLAIN=$pl timeout 600 python bench.py --ops-only --ops locate --steps 10 --no-e2e --no-cpu-baseline 2> $OUT/r3b_bench$pl.err ) > $OUT/r3b_bench$pl.json
tail -1 $OUT/r3b_bench$pl.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3b_bench$pl.json').read().strip().splitlines()[-1])
    for k,v in d['ops'].items(): print('plain=$pl',k,'ms',round(v['ms_per_step'],4),'kernel_ms',round(v['roofline']['kernel_ms'],4),'frac',round(v['roofline']['frac'],4),v.get('parity',{}).get('match'))
except Exception as e: print('ERR',e)
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_emit_contig -s 3 -c 1 -f -o $OUT/r3b_contig_prof \
  python bench.py --ops-only --ops rmdup --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > $OUT/r3b_ncu_contig.log 2>&1
BSK_LT_PLAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_locate_tile -s 3 -c 1 -f -o $OUT/r3b_locate_prof \
  python bench.py --ops-only --ops locate --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > $OUT/r3b_ncu_locate.log 2>&1
ls -la $OUT/r3b_*
