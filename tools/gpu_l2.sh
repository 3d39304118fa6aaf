#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/l_rmdup2.csv python tools/bench_ops.py --mib 256 --ops rmdup --steps 1 --warmup 1 > $OUT/l_rmdup2.log 2>&1
BSK_NO_CONTIG=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/l_rmdup3.csv python tools/bench_ops.py --mib 256 --ops rmdup --steps 1 --warmup 1 > $OUT/l_rmdup3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/l_translate2.csv python tools/bench_ops.py --mib 256 --ops translate --steps 1 --warmup 1 > $OUT/l_translate2.log 2>&1
