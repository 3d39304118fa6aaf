#!/bin/bash
# One GPU session for the record: gpu parity tests, bench (both arms), file -> file, then the profiler passes --
# launch list + ncu --set full of the dominant kernel of every workload of the bench.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/${TAG}_pytest.log
( timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2> $OUT/${TAG}_bench.err ) > $OUT/${TAG}_bench_ref.json
( timeout 900 python bench.py --steps 10 --warmup 3 2>> $OUT/${TAG}_bench.err ) > $OUT/${TAG}_bench.json
tail -5 $OUT/${TAG}_pytest.log; cut -c1-400 $OUT/${TAG}_bench.json; cut -c1-300 $OUT/${TAG}_bench_ref.json; tail -3 $OUT/${TAG}_bench.err
( timeout 600 python tools/bench_file.py --mib 4096 --threads 0 --cpu 2>> $OUT/${TAG}_bench.err ) > $OUT/${TAG}_file.jsonl; cat $OUT/${TAG}_file.jsonl
# ---- profiler passes (numbers printed under ncu are never bench values)
prof() {  # <name> <bench args> <kernel regex> <skip>
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_$1_launches.csv \
    python bench.py $2 --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > $OUT/${TAG}_$1_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o $OUT/${TAG}_$1_prof \
    python bench.py $2 --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > $OUT/${TAG}_$1_ncu_full.log 2>&1
}
prof seq "--ops none" '^k_fastq_inplace' 3
prof stats "--ops-only --ops stats" '^k_stats_tile' 3
prof stats_all "--ops-only --ops stats_all" '^k_stats_tile' 3
prof rmdup "--ops-only --ops rmdup" '^k_rmdup_tile$' 3
prof translate "--ops-only --ops translate" '^k_translate_tile$' 3
prof locate "--ops-only --ops locate" '^k_locate_tile$' 3
ls -la $OUT | grep ${TAG}_ | wc -l
