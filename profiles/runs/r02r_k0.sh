#!/bin/bash
# round 2, run r: K0 emit without atomics / valid predicate
set -x
mkdir -p gpurun_out
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 2>&1 | tail -3 | cut -c1-200
RDR_K0_MINB=5 timeout 300 python profiles/r02_check.py c2 ml145 2>&1 | tail -2 | cut -c1-200
RDR_K0_MINB=8 timeout 300 python profiles/r02_check.py c2 ml145 2>&1 | tail -2 | cut -c1-200
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "^(FAILED|E  +Assert|E  +assert|[0-9]+ (passed|failed))" | cut -c1-300 > gpurun_out/r02r_gpu_tests.txt
cat gpurun_out/r02r_gpu_tests.txt
