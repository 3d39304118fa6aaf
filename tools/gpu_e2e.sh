#!/bin/bash
run() { timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', 'ms_per_step %.4f e2e_GBps %.2f e2e_rec/s %.3e' % (d['ms_per_step'], d['e2e']['gb_per_s'], d['e2e']['value']))"; }
timeout 600 python -m pytest tests/test_fused_path.py tests/test_parity_rmdup.py tests/test_parity_seq.py -m gpu -x -q 2>&1 | tail -3
run --e2e-block-mib 1024
run --e2e-block-mib 256
run --e2e-block-mib 128
run --e2e-block-mib 64
run --e2e-block-mib 32
run --e2e-block-mib 16
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
