#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > $O/r02j_tests.txt 2>&1; tail -8 $O/r02j_tests.txt
for w in c5_manhattan_b8 c3_pearson_b8; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $O/r02j_$w.json 2>$O/r02j_$w.err
done
for w in c3_cosine_b8 c5_euclid_b8; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --path 2 > $O/r02j_${w}_sweep.json 2>$O/r02j_${w}_sweep.err
done
for w in c4_decay_mask_k100 c4_decay_mask_k100_clustered; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $O/r02j_$w.json 2>$O/r02j_$w.err
done
for f in $O/r02j_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value=%.1f e2e=%.1f sync=%.1f ms=%.3f frac=%.3f launches=%d unc=%s kern=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['gpu_launches'],d['config']['uncertified_steps'],d['roofline']['kernel']))
except Exception as e:
    print(sys.argv[1],"ERR",e)
PY
done
tail -n 3 $O/r02j_*.err
