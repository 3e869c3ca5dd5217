#!/bin/bash
# round 2, run m: the reference's test_slant goldens through the CUDA path + the whole GPU suite again
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "test_slant" 2>&1 | tail -15 > gpurun_out/r02m_slant.txt
cat gpurun_out/r02m_slant.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02m_gpu_tests.txt
cat gpurun_out/r02m_gpu_tests.txt
