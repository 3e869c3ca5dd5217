#!/bin/bash
# round 2, run ag: the reference's station-mode golden (test_intersect.py:104) through the product's tropo_delay on the device
set -x
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "gnss" 2>&1 | tail -15 | cut -c1-250
