#!/bin/bash
O=gpurun_out
python scratch/ab_old/bench.py --workload c4_decay_mask_k100 --steps 20 --warmup 3 --no-cpu-baseline > $O/r02k_old_c4.json 2>$O/r02k_old_c4.err
HDB_BENCH_MASK=random python bench.py --workload c4_decay_mask_k100 --steps 20 --warmup 3 --no-cpu-baseline > $O/r02k_new_c4_random.json 2>$O/r02k_new_c4_random.err
python bench.py --workload c4_decay_mask_k100 --steps 20 --warmup 3 --no-cpu-baseline > $O/r02k_new_c4_hash.json 2>$O/r02k_new_c4_hash.err
for f in $O/r02k_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value=%.1f e2e=%.1f ms=%.3f frac=%.3f launches=%d unc=%s"%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['gpu_launches'],d['config']['uncertified_steps']))
except Exception as e:
    print(sys.argv[1],"ERR",e)
PY
done
tail -n 3 $O/r02k_*.err
