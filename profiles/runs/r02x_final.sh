set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_tests.log
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02x_check.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_sample_stream --csv --log-file gpurun_out/r02x_k2_dram.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r02x_k2_dram_bench.log 2>&1
python profiles/ncu_traffic.py gpurun_out/r02x_k2_dram.csv 192000000 > gpurun_out/k2_traffic.json 2> gpurun_out/r02x_k2_traffic.err
cp gpurun_out/k2_traffic.json profiles/k2_traffic.json  # (the bench below reads it; same tree -> same kernel-source hash)
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02x_bench_reference.json 2> gpurun_out/r02x_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02x_launches_bench.log 2>&1
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/tmp_$3 -f ${@:4} > gpurun_out/r02x_ncu_$3.log 2>&1; bash profiles/summarize_ncu.sh gpurun_out/tmp_$3.ncu-rep gpurun_out/r02x_$3_ncu.txt; }
cap k_ray_integrate_poly 2 k3_poly_c2 python profiles/r02_check.py c2
cap k_ray_layers 2 k0_c2 python profiles/r02_check.py c2
cap k_ray_integrate_thin 2 k3_thin_ml145 python profiles/r02_check.py ml145
cap k_ray_layers 2 k0_ml145 python profiles/r02_check.py ml145
cap k_sample_stream 3 k2_stream python bench.py --steps 1 --warmup 1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "fused or staged or knife or all_nan or slant" > gpurun_out/r02x_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02x_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "staged or fused" > gpurun_out/r02x_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02x_racecheck.log
timeout 600 python bench.py --steps 5 --warmup 3 --config c5 > gpurun_out/r02x_bench_c5_n1.json 2> gpurun_out/r02x_bench_c5_n1.err
du -sh gpurun_out; tail -3 gpurun_out/r02x_tests.log; cat gpurun_out/r02x_check.log | cut -c1-200; tail -3 gpurun_out/r02x_bench.err; cut -c1-300 gpurun_out/r02x_bench_reference.json; tail -3 gpurun_out/r02x_memcheck.log gpurun_out/r02x_racecheck.log
