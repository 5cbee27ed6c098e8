#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r02q_tests.txt 2>&1; tail -5 $O/r02q_tests.txt
( time python bench.py > $O/r02q_bench_default.json 2> $O/r02q_bench_default.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02q_bench_default.json").read().strip().splitlines()[-1])
print("HEAD value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f launches=%d unc=%s parity=%s cpu=%s clocks=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['gpu_launches'],d['config']['uncertified_steps'],d.get('parity_check',{}).get('ok'),d.get('cpu_baseline',{}).get('value'),d['clocks']))
for e in d.get('extra',[]):
    if 'error' in e: print("   extra", e); continue
    print("   extra %-30s value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f unc=%s %s"%(e['workload'],e['value'],e['e2e']['value'],e['e2e']['sync_value'],e['ms_per_step'],e['roofline']['frac'] or 0,e['uncertified_steps'],e['roofline']['kernel']))
PY
tail -n 3 $O/r02q_bench_default.err
