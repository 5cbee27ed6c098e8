#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r02n_tests.txt 2>&1; tail -6 $O/r02n_tests.txt
python scratch/c4_probe.py > $O/r02n_c4_probe_staged.txt 2>&1; cat $O/r02n_c4_probe_staged.txt
for lim in 0 4096; do
  for w in c3_cosine_b1 c2_cosine_b1 c5_manhattan_b1; do
    HDB_SWEEP_STAGED=$lim python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline --extras none > $O/r02n_${w}_lim$lim.json 2>$O/r02n_${w}_lim$lim.err
  done
done
for f in $O/r02n_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f sm=%s unc=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['clocks']['sm_mhz'],d['config']['uncertified_steps']))
except Exception as e:
    print(sys.argv[1],"ERR",e)
PY
done
