#!/bin/bash
O=gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > $O/val_tests.txt 2>&1; tail -4 $O/val_tests.txt
( time python bench.py > $O/val_bench_default.json 2> $O/val_bench_default.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/val_bench_default.json").read().strip().splitlines()[-1])
print("HEAD value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f launches=%d unc=%s parity=%s cpu=%s clocks=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['gpu_launches'],d['config']['uncertified_steps'],d.get('parity_check',{}).get('ok'),d.get('cpu_baseline',{}).get('value'),d['clocks']))
for e in d.get('extra',[]):
    if 'error' in e or 'roofline' not in e: print("   extra", e); continue
    print("   extra %-30s value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f unc=%s %s"%(e['workload'],e['value'],e['e2e']['value'],e['e2e']['sync_value'],e['ms_per_step'],e['roofline']['frac'] or 0,e['uncertified_steps'],e['roofline']['kernel']))
PY
tail -n 3 $O/val_bench_default.err
for w in c2_cosine_b1 c5_euclid_b1 c3_pearson_b1 c5_euclid_b1024; do python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --extras none > $O/val_$w.json 2>$O/val_$w.err; done
for f in $O/val_c*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f unc=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['config']['uncertified_steps']))
except Exception as e:
    print(sys.argv[1],"ERR",e)
PY
done
# opt-in configuration (not part of the default run): queries wider than the store on the tensor path
HDB_TC_MIXED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wider_queries" > $O/val_tc_mixed.txt 2>&1; tail -3 $O/val_tc_mixed.txt
timeout 300 python scratch/pearson_tc_ab.py 10000000 768 1024 2>&1 | tail -1
