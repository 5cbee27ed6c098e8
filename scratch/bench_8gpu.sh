#!/bin/bash
# multi-GPU verification: NCCL/IPC test + the sharded benches.  usage: bash scratch/bench_8gpu.sh <n_gpus> <out_dir>
N=${1:-8}; OUT=${2:-gpurun_out/r02/g8}
mkdir -p $OUT
(timeout 240 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k nccl > $OUT/t_nccl.log 2>&1; echo rc=$? >> $OUT/t_nccl.log); tail -n 3 $OUT/t_nccl.log
run() { # workload exchange extra...
  w=$1; x=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --no-cpu-baseline --exchange $x "$@" > $OUT/bench${N}_${w}_$x.json 2> $OUT/bench${N}_${w}_$x.err || echo "FAILED $w $x"
  grep -h "^{" $OUT/bench${N}_${w}_$x.json | cut -c1-150
}
run c3_cosine_b1 peer
run c3_cosine_b1 nccl
run c4_decay_mask_k100 peer
run c5_hamming_b1 peer
run c3_cosine_b4096 peer --steps 12
