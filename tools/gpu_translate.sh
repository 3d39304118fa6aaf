#!/bin/bash
# translate iteration: parity, whole-step timing at 1 GiB, per-kernel launch list at 256 MiB
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_translate.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/bench_ops.py --mib 1024 --steps 5 --ops translate 2>&1 | cut -c1-330
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/l_translate8.csv python tools/bench_ops.py --mib 256 --ops translate --steps 1 --warmup 1 > $OUT/l_translate8.log 2>&1
grep -E "k_translate|k_emit\(" $OUT/l_translate8.csv | awk -F'","' '{print $5, $(NF)}' | head -6
