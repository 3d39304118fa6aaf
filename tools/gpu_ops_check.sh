#!/bin/bash
# whole GPU suite, then device-resident timings of the operators named in $1 (default: rmdup,fq2fa) with the CPU port beside
OUT=gpurun_out; mkdir -p $OUT
OPS=${1:-rmdup,fq2fa}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/bench_ops.py --mib 1024 --steps 5 --ops $OPS --cpu 2>&1 | tee $OUT/ops_check.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['op'], 'ms %.3f GB/s %.1f frac %.3f fused %d launches %d' % (d['ms_per_step'], d['gb_per_s'], d['whole_step_frac_of_hbm_peak'], d['fused_blocks'], d['gpu_launches_per_step']), 'cpu', d.get('cpu_port'))
"
