#!/bin/bash
export BSK_DEBUG=1
for v in 1 0; do for sz in 4194304 33554432 268435456; do echo "variant $v size $sz"; BSK_FQ_VARIANT=$v timeout 600 python tools/debug_fused.py $sz 1 2>&1 | tail -4; done; done
