#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "fused_equals_exact or row_order or multi_query" 2>&1 | tail -2
python -m pytest tests/test_gpu_hyperdb_shim.py tests/test_gpu_mutation.py -q -x 2>&1 | tail -2
python scratch/c4_probe.py > $O/r02r_c4_probe.txt 2>&1; cat $O/r02r_c4_probe.txt
