mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_locate_tile.py tests/test_parity_match.py tests/test_parity_rmdup.py tests/test_fullsize_gpu.py tests/test_exchange.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --ops-only --ops rmdup,locate --steps 5 --no-e2e > gpurun_out/r2t_ops.json 2> gpurun_out/r2t_ops.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_ops.json').read().strip().splitlines()[-1])
for k,v in d['ops'].items(): print(k, round(v['ms_per_step'],3), round(v['roofline']['kernel_ms'],3), round(v['roofline']['frac'],3), round(v['roofline']['whole_step_frac'],3), v['parity']['match'])
PY
tail -2 gpurun_out/r2t_ops.err
