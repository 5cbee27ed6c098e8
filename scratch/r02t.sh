#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r02t_tests.txt 2>&1; tail -5 $O/r02t_tests.txt
bash scratch/profile_round.sh $O/r02/prof > $O/r02t_prof.log 2>&1; tail -15 $O/r02t_prof.log
