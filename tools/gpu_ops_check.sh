#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for e in "" "BSK_NO_SEQ_TILE=1"; do echo "== $e"; env $e timeout 600 python tools/bench_ops.py --mib 1024 --steps 5 --ops seq_fastq_minlen,seq_fasta_rc,translate,subseq 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['op'], 'ms %.3f GB/s %.1f frac %.3f fused %d launches %d' % (d['ms_per_step'], d['gb_per_s'], d['whole_step_frac_of_hbm_peak'], d['fused_blocks'], d['gpu_launches_per_step']))
"; done
