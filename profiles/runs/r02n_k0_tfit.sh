#!/bin/bash
# round 2, run n: K0 with the layer tops as one polynomial in z (RDR_K0_MODE=poly, default) vs three iterates per layer (iter)
set -x
mkdir -p gpurun_out
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02n_check_poly.log 2>&1
RDR_K0_MODE=iter timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02n_check_iter.log 2>&1
tail -4 gpurun_out/r02n_check_poly.log; tail -4 gpurun_out/r02n_check_iter.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02n_gpu_tests.txt
cat gpurun_out/r02n_gpu_tests.txt
