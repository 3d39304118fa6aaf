#!/bin/bash
run() { env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', 'ms_per_step %.4f kernel_ms %.4f frac %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"; }
run BSK_FQ_VARIANT=3
run BSK_FQ_VARIANT=6
run BSK_FQ_VARIANT=6 BSK_FQ_EARLY=1
run BSK_FQ_VARIANT=3
run BSK_FQ_VARIANT=6
