#!/bin/bash
# round 2, run w: K0 row update by lane 0 alone, predicated; racecheck on K0 + staged thin kernel
set -x
mkdir -p gpurun_out
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 2>&1 | tail -3 | cut -c1-130
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "staged" > gpurun_out/r02w_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02w_racecheck.log
tail -4 gpurun_out/r02w_racecheck.log
