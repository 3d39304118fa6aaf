#!/bin/bash
# one optimisation iteration on the GPU: tile-path parity tests, seq bench, ncu full capture of the top kernel
TAG=${1:-i}; KRE=${2:-k_fastq_inplace}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fused_path.py tests/test_golden.py tests/test_properties_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --ops none > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-1200 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 2 --warmup 3 --block-mib 256 --no-e2e --no-cpu-baseline --no-parity --ops none > $OUT/${TAG}_ncu_full.log 2>&1
