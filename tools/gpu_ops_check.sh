#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_translate.py tests/test_golden.py tests/test_cli.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/bench_ops.py --mib 1024 --steps 5 --ops translate 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['op'], 'ms %.3f GB/s %.1f frac %.3f fused %d launches %d' % (d['ms_per_step'], d['gb_per_s'], d['whole_step_frac_of_hbm_peak'], d['fused_blocks'], d['gpu_launches_per_step']))
"
for op in translate; do timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/l_${op}6.csv python tools/bench_ops.py --mib 256 --ops $op --steps 1 --warmup 1 > $OUT/l_${op}6.log 2>&1; done
