for c in 0 4 1; do BSK_FQ_COPYONLY=1 BSK_FQ_SHAPE=$c timeout 300 python bench.py --steps 10 --warmup 3 --ops none --no-cpu-baseline --no-e2e --no-parity 2>&1 | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print('copy-only shape $c', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'ms_per_step', round(d['ms_per_step'],4))
except Exception as e: print('fail', t[-300:])"; done
