set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_tests.log
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02k_check.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02k_bench_reference.json 2> gpurun_out/r02k_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02k_launches_bench.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_sample_stream --csv --log-file gpurun_out/r02k_k2_dram.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r02k_k2_dram_bench.log 2>&1
python profiles/ncu_traffic.py gpurun_out/r02k_k2_dram.csv 192000000 > gpurun_out/k2_traffic.json 2> gpurun_out/r02k_k2_traffic.err
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/tmp_$3 -f ${@:4} > gpurun_out/r02k_ncu_$3.log 2>&1; bash profiles/summarize_ncu.sh gpurun_out/tmp_$3.ncu-rep gpurun_out/r02k_$3_ncu.txt; }
cap k_ray_integrate_poly 2 k3_poly_c2 python profiles/r02_check.py c2
cap k_ray_layers 2 k0_c2 python profiles/r02_check.py c2
cap k_ray_integrate_thin 2 k3_thin_ml145 python profiles/r02_check.py ml145
cap k_ray_layers 2 k0_ml145 python profiles/r02_check.py ml145
cap k_sample_stream 3 k2_stream python bench.py --steps 1 --warmup 1
RDR_K3_UNIFIED=1 timeout 200 python profiles/r02_check.py c2 ml145 > gpurun_out/r02k_unified.log 2>&1
RDR_K3_THIN_STAGE=0 timeout 200 python profiles/r02_check.py ml145 > gpurun_out/r02k_unstaged.log 2>&1
( cd profiles/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_regs dfma_regs.cu && /tmp/dfma_regs > ../../gpurun_out/r02_dfma_regs.txt 2>&1; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_latency dfma_latency.cu && /tmp/dfma_latency > ../../gpurun_out/r02_dfma_latency.txt 2>&1 )
du -sh gpurun_out; tail -3 gpurun_out/r02k_tests.log; cat gpurun_out/r02k_check.log; tail -3 gpurun_out/r02k_bench.err; cut -c1-300 gpurun_out/r02k_bench_reference.json
