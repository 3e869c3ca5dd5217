#!/bin/bash
# round 2, run u: compute-sanitizer (memcheck, then racecheck on the staged thin-layer kernel) over small GPU tests
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "fused or staged or knife or all_nan or slant" > gpurun_out/r02u_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02u_memcheck.log
tail -8 gpurun_out/r02u_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "staged" > gpurun_out/r02u_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02u_racecheck.log
tail -8 gpurun_out/r02u_racecheck.log
