#!/bin/bash
# ncu evidence for profiles/: every capture only after the same command exited 0 without ncu (B200_PROFILING.md).
# usage (on the GPU box, from the repo root): bash scratch/profile_round.sh <out_dir>
set -u
OUT=${1:-gpurun_out/r02/prof}
mkdir -p $OUT
B="python bench.py --no-cpu-baseline --extras none"
KERN='regex:sweep|finalize|prep_query|merge|batched_tc|bucket_records|sample_threshold|queries_to_half|exchange'
run_plain() { # workload steps
  $B --workload $1 --steps $2 --warmup 3 > $OUT/plain_$1.json 2> $OUT/plain_$1.err
}
launch_list() { # workload steps count
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c $3 --csv --log-file $OUT/launches_$1.csv \
      $B --workload $1 --steps $2 --warmup 3 > $OUT/ncu_list_$1.log 2>&1
}
full() { # workload kernel-regex skip tag
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -o $OUT/$4 -f \
      $B --workload $1 --steps 3 --warmup 3 > $OUT/ncu_full_$4.log 2>&1
}
run_plain c3_cosine_b1 5 && launch_list c3_cosine_b1 5 120 && full c3_cosine_b1 'sweep_kernel' 4 sweep_c3
run_plain c5_hamming_b1 5 && launch_list c5_hamming_b1 5 120 && full c5_hamming_b1 'sweep_hamming' 4 ham_c5
run_plain c4_decay_mask_k100_clustered 3 && launch_list c4_decay_mask_k100_clustered 3 60 && full c4_decay_mask_k100_clustered 'sweep_staged' 3 staged_c4
run_plain c5_manhattan_b8 3 && launch_list c5_manhattan_b8 3 60
run_plain c3_cosine_b64 3 && full c3_cosine_b64 'batched_tc_kernel' 5 tc_b64
ls -la $OUT
