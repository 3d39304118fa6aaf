#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_stats_tile.py tests/test_parity_seq.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/bench_ops.py --mib 1024 --ops stats_fasta,stats_fastq --steps 10 2>&1 | cut -c1-900
BSK_NO_STATS_TILE=1 timeout 900 python tools/bench_ops.py --mib 1024 --ops stats_fasta,stats_fastq --steps 5 2>&1 | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stats_tile -s 3 -c 1 -f -o $OUT/stats_prof \
  python tools/bench_ops.py --mib 256 --ops stats_fasta --steps 2 > $OUT/stats_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stats_tile -s 3 -c 1 -f -o $OUT/statsq_prof \
  python tools/bench_ops.py --mib 256 --ops stats_fastq --steps 2 > $OUT/statsq_ncu.log 2>&1
