#!/bin/bash
# A/B of the single-query sweep before / after the multi-query rewrite, same box, alternating
set -x
O=gpurun_out
for i in 1 2; do
  python scratch/ab_old/bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/r02d_old_c3b1_$i.json 2>$O/r02d_old_$i.err
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/r02d_new_c3b1_$i.json 2>$O/r02d_new_$i.err
done
python scratch/ab_old/bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02d_old_c3b1_s20.json 2>>$O/r02d_old_1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02d_new_c3b1_s20.json 2>>$O/r02d_new_1.err
for w in c5_manhattan_b8 c5_euclid_b8 c5_hamming_b8 c3_cosine_b8 c3_pearson_b8 c5_hamming_b1; do
  python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline > $O/r02d_$w.json 2>$O/r02d_$w.err
done
python -m pytest tests/test_gpu_sharded.py -q --timeout 600 > $O/r02d_sharded_tests.txt 2>&1
tail -5 $O/r02d_sharded_tests.txt
grep -H -o '"value": [0-9.]*\|"sm_mhz": [0-9.]*\|"frac": [0-9.]*' $O/r02d_*.json | paste - - - - -
