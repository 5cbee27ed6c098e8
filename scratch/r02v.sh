#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "multi_query or row_order or fused_equals" 2>&1 | tail -3
python -m pytest tests/test_gpu_fullsize.py -q -x -k "b8 or B8 or manhattan_b8 or dot_b8 or hamming_b8" 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --extras none --steps 20 --warmup 3"
$B --workload c5_manhattan_b8 > $O/r02v_manh_b8.json 2>$O/r02v_1.err
$B --workload c3_cosine_b8 --path 2 > $O/r02v_cos_b8_sweep.json 2>$O/r02v_2.err
$B --workload c3_pearson_b8 > $O/r02v_pearson_b8.json 2>$O/r02v_3.err
$B --workload c5_euclid_b8 --path 2 > $O/r02v_euclid_b8_sweep.json 2>$O/r02v_4.err
$B --workload c5_manhattan_b8 --max-group 4 > $O/r02v_manh_b8_g4.json 2>$O/r02v_5.err
HDB_SWEEP_STAGED_MQ=0 $B --workload c5_manhattan_b8 > $O/r02v_manh_b8_reg.json 2>$O/r02v_6.err
for f in $O/r02v_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value=%.1f e2e=%.1f ms=%.4f frac=%.3f unc=%s"%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['config']['uncertified_steps']))
except Exception as e:
    print(sys.argv[1],"ERR",e)
PY
done
tail -n 3 $O/r02v_*.err | tail -20
