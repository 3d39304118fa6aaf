mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_locate_tile.py tests/test_parity_match.py tests/test_parity_rmdup.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_prof.sh r2p_locate locate k_locate_tile 1024 | cut -c1-1200
timeout 600 python bench.py --ops-only --ops rmdup --steps 5 --no-e2e > gpurun_out/r2p_rmdup.json 2> gpurun_out/r2p_rmdup.err; cut -c1-900 gpurun_out/r2p_rmdup.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_rmdup_tile$ -s 3 -c 1 -f -o gpurun_out/r2p_rmdup_prof python bench.py --ops-only --ops rmdup --steps 2 --warmup 3 --no-e2e --no-parity > gpurun_out/r2p_rmdup_ncu.log 2>&1
