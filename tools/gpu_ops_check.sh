#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/bench_ops.py --mib 1024 --ops rmdup,stats_fastq --steps 5 2>&1 | cut -c1-330
BSK_NO_RMDUP_TILE=1 timeout 900 python tools/bench_ops.py --mib 1024 --ops rmdup --steps 5 2>&1 | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/l_rmdup4.csv python tools/bench_ops.py --mib 256 --ops rmdup --steps 1 --warmup 1 > $OUT/l_rmdup4.log 2>&1
