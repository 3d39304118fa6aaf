#!/bin/bash
# One workload under the profiler: timing line, ncu launch list, ncu --set full of its dominant kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_prof.sh <tag> <workload> <kernel-regex> [block-mib]
set -u
TAG=$1; WL=$2; KRE=$3; MIB=${4:-1024}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python bench.py --ops-only --ops $WL --steps 5 --block-mib $MIB 2> $OUT/${TAG}_bench.err ) > $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --ops-only --ops $WL --steps 2 --warmup 3 --block-mib $MIB --no-e2e --no-parity > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/${TAG}_prof \
  python bench.py --ops-only --ops $WL --steps 2 --warmup 3 --block-mib $MIB --no-e2e --no-parity > $OUT/${TAG}_ncu_full.log 2>&1
cut -c1-1500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
