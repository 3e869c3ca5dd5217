#!/bin/bash
# round 2, run ac: upper `.all()` clamp of the last sample (delay.py:310-311)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "^(FAILED|E  +Assert|E  +assert|E  |[0-9]+ (passed|failed))" | cut -c1-300 | head -40
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 2>&1 | tail -3 | cut -c1-200
