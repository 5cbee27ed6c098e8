#!/bin/bash
O=gpurun_out
scratch/micro/inner > $O/r02g_inner.txt 2>&1; cat $O/r02g_inner.txt
python scratch/ab_sweep.py scratch/ab_old/local-hyperdb_b200 local-hyperdb_b200 > $O/r02g_ab_sweep.txt 2>&1; cat $O/r02g_ab_sweep.txt
python -m pytest tests -m gpu -q -x --timeout 900 > $O/r02g_tests.txt 2>&1; tail -15 $O/r02g_tests.txt
