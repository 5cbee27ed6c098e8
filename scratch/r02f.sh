#!/bin/bash
O=gpurun_out
python scratch/ab_sweep.py scratch/ab_old/local-hyperdb_b200 local-hyperdb_b200 > $O/r02f_ab_sweep.txt 2>&1; cat $O/r02f_ab_sweep.txt
for w in c5_manhattan_b8 c5_hamming_b8 c3_pearson_b8; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $O/r02f_$w.json 2>$O/r02f_$w.err
done
for w in c3_cosine_b8 c5_euclid_b8; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --path 2 > $O/r02f_${w}_sweep.json 2>$O/r02f_${w}_sweep.err
done
python bench.py --workload c5_manhattan_b8 --steps 20 --warmup 3 --no-cpu-baseline --max-group 4 > $O/r02f_c5_manhattan_b8_g4.json 2>$O/r02f_g4.err
python bench.py --workload c3_cosine_b4096 --steps 6 --warmup 3 --no-cpu-baseline > $O/r02f_c3_cosine_b4096.json 2>$O/r02f_b4096.err
for f in $O/r02f_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value=%.1f e2e=%.1f sync=%.1f ms=%.3f frac=%.3f launches=%d unc=%s kern=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['gpu_launches'],d['config']['uncertified_steps'],d['roofline']['kernel']))
except Exception as e:
    print(sys.argv[1],"ERR",e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02f_launches_c5_hamming_b8.csv python bench.py --workload c5_hamming_b8 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02f_ncu_ham.log 2>&1
tail -3 $O/*.err | tail -30
