mkdir -p gpurun_out
( timeout 400 python bench.py --steps 10 --warmup 3 --ops none 2> gpurun_out/r2x_bench.err ) > gpurun_out/r2x_bench.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print('seq', d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['l2_policy'][-60:], d['parity']['match'])
PY
tail -2 gpurun_out/r2x_bench.err
bash tools/gpu_sanitize.sh
