#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_seq.py tests/test_parity_fq2fa.py tests/test_parity_match.py tests/test_parity_translate.py tests/test_parity_raw_ops.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python tools/bench_ops.py --mib 1024 --steps 5 --ops subseq,grep_id,fq2fa,seq_fastq_minlen,seq_fasta_rc,seq_fasta_reads 2>&1 | tee $OUT/r3g_ops.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['op'], 'ms %.3f GB/s %.1f frac %.3f fused %d launches %d' % (d['ms_per_step'], d['gb_per_s'], d['whole_step_frac_of_hbm_peak'], d['fused_blocks'], d['gpu_launches_per_step']))
"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_emit -s 1 -c 1 -f -o $OUT/r3g_emit_prof \
  python tools/bench_ops.py --mib 1024 --steps 1 --warmup 1 --ops subseq > $OUT/r3g_ncu_emit.log 2>&1
for op in subseq seq_fasta_rc; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r3g_launches_$op.csv \
  python tools/bench_ops.py --mib 1024 --steps 2 --warmup 1 --ops $op > $OUT/r3g_ncu_$op.log 2>&1
done
python - <<'PY'
import csv,collections
for f in ('r3g_launches_subseq.csv','r3g_launches_seq_fasta_rc.csv'):
    try:
        rows=[r for r in csv.reader(l for l in open('gpurun_out/'+f) if l.startswith('"'))]
        h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
        agg=collections.defaultdict(list)
        for r in rows[1:]: agg[r[ki][:40]].append(float(r[vi].replace(',','')))
        print(f)
        for k,v in sorted(agg.items(), key=lambda x:-sum(x[1]))[:9]: print('   %-42s n=%3d avg %.1f us'%(k,len(v),sum(v)/len(v)/1000))
    except Exception as e: print(f,'ERR',e)
PY
