#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_seq.py tests/test_fused_path.py tests/test_file_entry.py tests/test_abi.py tests/test_exchange.py -m gpu -x -q 2>&1 | tail -2
for k in 1 2; do
( timeout 600 python bench.py --ops none --steps 10 --warmup 3 2> $OUT/r3i_bench.err ) > $OUT/r3i_bench$k.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r3i_bench$k.json').read().strip().splitlines()[-1])
print('seq value %.4g ms %.4f e2e %.4g rec/s %.2f GB/s frac_pcie %.3f probe %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['gb_per_s'],d['e2e']['frac_of_pcie_probe'],d['pcie_roofline']['bidir_gbs_per_direction']))
PY
done
( timeout 600 python tools/bench_file.py --mib 4096 --threads 0 2>> $OUT/r3i_bench.err ) > $OUT/r3i_file.jsonl; cut -c1-300 $OUT/r3i_file.jsonl
( timeout 600 python bench.py --ops-only --ops stats,rmdup,translate --steps 5 --no-cpu-baseline 2>> $OUT/r3i_bench.err ) > $OUT/r3i_ops.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3i_ops.json').read().strip().splitlines()[-1])
for k,v in d['ops'].items(): print(k,'ms %.4f e2e GB/s %s'%(v['ms_per_step'],(v.get('e2e') or {}).get('gb_per_s')))
PY
tail -2 $OUT/r3i_bench.err
