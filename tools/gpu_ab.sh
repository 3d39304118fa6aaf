mkdir -p gpurun_out
BSK_FQ_CTAS=4 timeout 600 python -m pytest tests/test_fused_path.py tests/test_properties_gpu.py -m gpu -x -q 2>&1 | tail -2
for c in 3 4 3 4; do BSK_FQ_CTAS=$c timeout 300 python bench.py --steps 10 --warmup 3 --ops none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ctas $c', 'ms_per_step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"; done
