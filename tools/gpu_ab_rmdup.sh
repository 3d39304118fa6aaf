#!/bin/bash
# rmdup: parity subset, bench line, per-kernel times (ncu launch list) -- one gpurun call per A/B step
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_rmdup.py tests/test_formatter_seams.py tests/test_parity_match.py -m gpu -x -q 2>&1 | tail -2
( timeout 600 python bench.py --ops-only --ops rmdup --steps 10 --no-e2e --no-cpu-baseline 2> $OUT/ab_rmdup.err ) > $OUT/ab_rmdup.json
tail -1 $OUT/ab_rmdup.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/ab_rmdup_launches.csv \
  python bench.py --ops-only --ops rmdup --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > $OUT/ab_rmdup_ncu.log 2>&1
python - <<'PY'
import json,csv,collections
d=json.loads(open('gpurun_out/ab_rmdup.json').read().strip().splitlines()[-1])
for k,v in d['ops'].items(): print(k,'ms',round(v['ms_per_step'],4),'kernel_ms',round(v['roofline']['kernel_ms'],4),v.get('parity',{}).get('match'))
rows=[r for r in csv.reader(l for l in open('gpurun_out/ab_rmdup_launches.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]: agg[r[ki][:40]].append(float(r[vi].replace(',','')))
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1]))[:6]: print('   %-42s n=%3d avg %.1f us'%(k,len(v),sum(v)/len(v)/1000))
PY
