#!/bin/bash
# ncu launch list + one full capture of the top kernel (1 GPU). usage: bash tools/gpu_ncu.sh <tag> <kernel-regex> [bench args]
TAG=${1:-p}; KRE=${2:-k_fastq_inplace}; shift; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --block-mib 256 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 2 --warmup 3 --block-mib 256 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_launch.log $OUT/${TAG}_ncu_full.log; ls -la $OUT
