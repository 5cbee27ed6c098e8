#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x --timeout 600 -k "not C4" > $O/r02p_tests.txt 2>&1; tail -4 $O/r02p_tests.txt
echo "--- C4 dense staged+fhfma"; python scratch/staged_probe.py
echo "--- C4 dense staged, no fhfma"; HDB_NO_FHFMA=1 python scratch/staged_probe.py
echo "--- C4 dense register+fhfma"; HDB_SWEEP_STAGED=0 python scratch/staged_probe.py
echo "--- C3 register+fhfma"; python scratch/staged_probe.py 10000000 768
echo "--- C3 register no fhfma"; HDB_NO_FHFMA=1 python scratch/staged_probe.py 10000000 768
echo "--- C3 staged+fhfma"; HDB_SWEEP_STAGED=2048 python scratch/staged_probe.py 10000000 768
python bench.py --workload c3_cosine_b8 --steps 20 --warmup 3 --no-cpu-baseline --path 2 > $O/r02p_c3_cosine_b8_sweep.json 2>$O/r02p_b8.err
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --extras none > $O/r02p_c3_cosine_b1.json 2>$O/r02p_b1.err
for f in $O/r02p_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f sm=%s unc=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['clocks']['sm_mhz'],d['config']['uncertified_steps']))
except Exception as e:
    print(sys.argv[1],"ERR",e)
PY
done
