#!/bin/bash
# round 2, run ai: after the source was split into per-kernel fragments (identical SASS: cuobjdump -sass md5 unchanged) the kernel-source
# hash moved -> K2 DRAM traffic re-captured for it; GPU suite + one bench line on the committed tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_sample_stream --csv --log-file gpurun_out/r02ai_k2_dram.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r02ai_k2_dram_bench.log 2>&1
python profiles/ncu_traffic.py gpurun_out/r02ai_k2_dram.csv 192000000 > gpurun_out/k2_traffic.json 2> gpurun_out/r02ai_k2_traffic.err
cp gpurun_out/k2_traffic.json profiles/k2_traffic.json
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02ai_bench.json 2> gpurun_out/r02ai_bench.err; cut -c1-230 gpurun_out/r02ai_bench.json
