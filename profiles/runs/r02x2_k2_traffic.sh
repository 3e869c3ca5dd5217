#!/bin/bash
# round 2: K2 DRAM traffic of the final tree (comment-only source changes after r02x_final.sh moved the kernel-source hash) + one bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_sample_stream --csv --log-file gpurun_out/r02x2_k2_dram.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r02x2_k2_dram_bench.log 2>&1
python profiles/ncu_traffic.py gpurun_out/r02x2_k2_dram.csv 192000000 > gpurun_out/k2_traffic.json 2> gpurun_out/r02x2_k2_traffic.err
cp gpurun_out/k2_traffic.json profiles/k2_traffic.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02x2_bench.json 2> gpurun_out/r02x2_bench.err
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
cut -c1-400 gpurun_out/r02x2_bench.json
