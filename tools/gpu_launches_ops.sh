#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for op in translate rmdup locate stats_fastq; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/l_${op}.csv python tools/bench_ops.py --mib 256 --ops $op --steps 1 --warmup 1 > $OUT/l_${op}.log 2>&1
done
timeout 600 python tools/bench_ops.py --mib 1024 --ops stats_fastq --steps 10 | cut -c1-400
