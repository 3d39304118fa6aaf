mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_translate_tile.py tests/test_parity_translate.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --ops-only --ops translate --steps 5 --no-e2e > gpurun_out/r2w_tr.json 2> gpurun_out/r2w_tr.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_tr.json').read().strip().splitlines()[-1])
for k,v in d['ops'].items(): print(k, round(v['ms_per_step'],3), round(v['roofline']['kernel_ms'],4), round(v['roofline']['frac'],3), v['roofline']['stage_ms'], v['parity']['match'])
PY
tail -2 gpurun_out/r2w_tr.err
