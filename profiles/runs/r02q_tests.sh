#!/bin/bash
# round 2, run q: GPU suite with the new defaults (K0 poly, 24 km spans, thick tail absorbed)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "^(FAILED|E  +Assert|E  +assert|[0-9]+ (passed|failed))" | cut -c1-300 > gpurun_out/r02q_gpu_tests.txt
cat gpurun_out/r02q_gpu_tests.txt
