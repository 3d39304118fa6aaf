mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_path.py tests/test_stats_tile.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for g in 8 4; do BSK_FQ_GROUP=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --ops none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('group $g', 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"; done
timeout 600 python bench.py --ops-only --ops stats,stats_all --steps 10 --no-e2e > gpurun_out/r2o_stats.json 2> gpurun_out/r2o_stats.err; python -c "
import json; d=json.load(open('gpurun_out/r2o_stats.json'))
for k,v in d['ops'].items(): print(k, v['ms_per_step'], v['roofline']['kernel_ms'], v['roofline']['frac'], v['parity'])
"; tail -2 gpurun_out/r2o_stats.err
