#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/bench_ops.py --mib 512 --ops rmdup,translate,locate --steps 5 2>&1 | cut -c1-330
BSK_NO_CONTIG=1 timeout 900 python tools/bench_ops.py --mib 512 --ops rmdup --steps 5 2>&1 | cut -c1-330
