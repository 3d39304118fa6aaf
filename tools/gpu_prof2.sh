#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in 0 1; do
BSK_FQ_VARIANT=$v timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fastq_inplace -s 3 -c 1 -f -o $OUT/v${v}_prof \
  python bench.py --steps 2 --warmup 3 --block-mib 256 --no-e2e --no-cpu-baseline > $OUT/v${v}_ncu_full.log 2>&1
done
ls -la $OUT | tail -5
