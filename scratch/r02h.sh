#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_hyperdb_shim.py tests/test_gpu_sharded.py -q -x --timeout 900 > $O/r02h_tests.txt 2>&1; tail -25 $O/r02h_tests.txt
B="python bench.py --no-cpu-baseline --extras none"
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 1 -o $O/r02h_mq_manh -f $B --workload c5_manhattan_b8 --steps 3 --warmup 3 > $O/r02h_ncu_manh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_hamming -s 4 -c 1 -o $O/r02h_mq_ham -f $B --workload c5_hamming_b8 --steps 3 --warmup 3 > $O/r02h_ncu_ham.log 2>&1
ls -la $O/*.ncu-rep | tail -3
