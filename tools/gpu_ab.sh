mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_rmdup.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -2
for mib in 0 32 128 0; do timeout 300 python bench.py --steps 10 --warmup 3 --ops none --no-cpu-baseline --e2e-block-mib $mib 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('e2e block $mib MiB', round(d['e2e']['gb_per_s'],2), 'probe', round(d['pcie_roofline']['bidir_gbs_per_direction'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],4))"; done
timeout 600 python bench.py --ops-only --ops rmdup --steps 5 > gpurun_out/r2r_rmdup.json 2> gpurun_out/r2r_rmdup.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_rmdup.json').read().strip().splitlines()[-1])
for k,v in d['ops'].items(): print(k, round(v['ms_per_step'],3), round(v['roofline']['kernel_ms'],3), round(v['roofline']['frac'],3), round(v['roofline']['whole_step_frac'],3), round(v['e2e']['gb_per_s'],1), v['parity']['match'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2r_rmdup_launches.csv python bench.py --ops-only --ops rmdup --steps 2 --warmup 3 --no-e2e --no-parity > /dev/null 2>&1
