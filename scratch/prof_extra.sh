#!/bin/bash
OUT=gpurun_out/r02/prof2; mkdir -p $OUT
B="python bench.py --no-cpu-baseline --extras none"
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 1 -o $OUT/mq_manh_final -f $B --workload c5_manhattan_b8 --steps 3 --warmup 3 > $OUT/ncu_mq.log 2>&1
ncu --set full --clock-control none -k regex:finalize_kernel -s 4 -c 1 -o $OUT/finalize_c3 -f $B --workload c3_cosine_b1 --steps 3 --warmup 3 > $OUT/ncu_fin.log 2>&1
ncu --set full --clock-control none -k regex:prep_query_kernel -s 4 -c 1 -o $OUT/prep_c3 -f $B --workload c3_cosine_b1 --steps 3 --warmup 3 > $OUT/ncu_prep.log 2>&1
ls -la $OUT
