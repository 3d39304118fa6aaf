#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_rmdup.py tests/test_exchange.py tests/test_formatter_seams.py -m gpu -x -q 2>&1 | tail -3
( timeout 600 python bench.py --ops-only --ops rmdup --steps 10 --no-e2e --no-cpu-baseline 2> $OUT/r3l_bench.err ) > $OUT/r3l_bench.json
tail -1 $OUT/r3l_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/r3l_launches_rmdup.csv \
  python bench.py --ops-only --ops rmdup --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > $OUT/r3l_ncu_rmdup.log 2>&1
python - <<'PY'
import json,csv,collections
for f in ('r3l_bench.json',):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        for k,v in d['ops'].items(): print(f,k,'ms',round(v['ms_per_step'],4),'kernel_ms',round(v['roofline']['kernel_ms'],4),'frac',round(v['roofline']['frac'],4),v.get('parity',{}).get('match'))
    except Exception as e: print(f,'ERR',e)
for f in ('r3l_launches_rmdup.csv',):
    try:
        rows=[r for r in csv.reader(l for l in open('gpurun_out/'+f) if l.startswith('"'))]
        h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
        agg=collections.defaultdict(list)
        for r in rows[1:]: agg[r[ki][:40]].append(float(r[vi].replace(',','')))
        print(f)
        for k,v in sorted(agg.items(), key=lambda x:-sum(x[1]))[:12]: print('   %-42s n=%3d avg %.1f us'%(k,len(v),sum(v)/len(v)/1000))
    except Exception as e: print(f,'ERR',e)
PY
