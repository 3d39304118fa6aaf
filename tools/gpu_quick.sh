#!/bin/bash
# quick GPU check of the seq hot path: parity on the debug input, fused-path tests, short bench
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-q}
timeout 300 python tools/debug_fused.py 4194304 2
timeout 900 python -m pytest tests/test_fused_path.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
