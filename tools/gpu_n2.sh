#!/bin/bash
# 2 GPUs of one box: NCCL exchange tests + bench with the exchange steps inside the timed region
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm --format=csv > $OUT/r2_n2_smi.txt 2>&1
( timeout 600 python -m pytest tests/test_dist_nccl.py tests/test_exchange.py -m gpu -q 2>&1 | tail -3 ) > $OUT/r2_n2_pytest.log; cat $OUT/r2_n2_pytest.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> $OUT/r2_n2_bench.err ) > $OUT/r2_n2_bench.json
tail -2 $OUT/r2_n2_bench.err; cut -c1-300 $OUT/r2_n2_bench.json
