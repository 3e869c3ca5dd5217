#!/bin/bash
# round 2, last call: what the driver runs at round end, on the committed tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02ah_bench.json 2> gpurun_out/r02ah_bench.err; cut -c1-260 gpurun_out/r02ah_bench.json; tail -2 gpurun_out/r02ah_bench.err
