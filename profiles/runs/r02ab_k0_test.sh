#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "k0_layer_top" 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
