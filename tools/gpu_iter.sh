#!/bin/bash
# one optimisation iteration on the GPU: fused-path parity tests, variant sweep, bench, ncu full capture of the top kernel
TAG=${1:-i}; KRE=${2:-k_fastq_inplace}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fused_path.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do for g in 8 16; do BSK_FQ_VARIANT=$v BSK_FQ_GROUP=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v group $g', 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"; done; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 2 --warmup 3 --block-mib 256 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
