#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_small.py > $OUT/san_$tool.log 2>&1
  grep -E "ok$|MISMATCH|SUMMARY|Error|error" $OUT/san_$tool.log | head -20
done
