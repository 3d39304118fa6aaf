#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_stats_tile.py tests/test_parity_seq.py tests/test_fullsize_gpu.py -k "stats" -m gpu -x -q 2>&1 | tail -2
( timeout 600 python bench.py --ops-only --ops stats,stats_all --steps 10 --no-e2e --no-cpu-baseline 2> $OUT/r3k_bench.err ) > $OUT/r3k_bench.json
tail -1 $OUT/r3k_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stats_tile -s 3 -c 1 -f -o $OUT/r3k_sa_prof \
  python bench.py --ops-only --ops stats_all --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > $OUT/r3k_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3k_bench.json').read().strip().splitlines()[-1])
for k,v in d['ops'].items(): print(k,'ms',round(v['ms_per_step'],4),'kernel_ms',round(v['roofline']['kernel_ms'],4),'frac',round(v['roofline']['frac'],4),v.get('parity',{}).get('match'))
PY
