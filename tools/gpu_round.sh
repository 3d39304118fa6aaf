#!/bin/bash
# One GPU session: gpu parity tests, bench, ncu launch list, ncu full capture of the top kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag> [kernel-regex]
set -u
TAG=${1:-r1}
KRE=${2:-k_fastq_inplace}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/${TAG}_pytest.log
( timeout 600 python bench.py --steps 10 --warmup 3 2> $OUT/${TAG}_bench.err ) > $OUT/${TAG}_bench.json
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/${TAG}_bench.err ) > $OUT/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --block-mib 256 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 2 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 2 --warmup 3 --block-mib 256 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -5 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_bench.json; cat $OUT/${TAG}_bench_ref.json; tail -3 $OUT/${TAG}_bench.err
( timeout 1200 python tools/bench_ops.py --mib 1024 --steps 5 --cpu 2>> $OUT/${TAG}_bench.err ) > $OUT/${TAG}_ops.jsonl
cut -c1-260 $OUT/${TAG}_ops.jsonl
