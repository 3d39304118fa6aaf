#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== plain"; timeout 300 python tools/debug_fused.py 4194304 2
echo "== ncu time"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file $OUT/dbg_launch.csv python tools/debug_fused.py 4194304 2
echo "== ncu cache-control none"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 20 --csv --log-file $OUT/dbg_launch2.csv python tools/debug_fused.py 4194304 2
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck python tools/debug_fused.py 1048576 1 2>&1 | tail -30
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck python tools/debug_fused.py 262144 1 2>&1 | tail -40
echo "== initcheck"; timeout 600 compute-sanitizer --tool initcheck python tools/debug_fused.py 1048576 1 2>&1 | tail -40
echo "== synccheck"; timeout 600 compute-sanitizer --tool synccheck python tools/debug_fused.py 262144 1 2>&1 | tail -20
