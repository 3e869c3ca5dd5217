set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_tests.log
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02j_check.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02j_bench_reference.json 2> gpurun_out/r02j_bench_reference.err
# launch list of the bench command (per-launch times, cold-cache and serialised: the SHARE of the step is what must agree)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02j_launches_bench.log 2>&1
# DRAM traffic of the K2 launch
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_sample_stream --csv --log-file gpurun_out/r02j_k2_dram.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r02j_k2_dram_bench.log 2>&1
python profiles/ncu_traffic.py gpurun_out/r02j_k2_dram.csv 192000000 > gpurun_out/k2_traffic.json 2> gpurun_out/r02j_k2_traffic.err
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_integrate_poly -s 2 -c 1 -o gpurun_out/r02j_k3_poly_c2 -f python profiles/r02_check.py c2 > gpurun_out/r02j_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_layers -s 2 -c 1 -o gpurun_out/r02j_k0_c2 -f python profiles/r02_check.py c2 > gpurun_out/r02j_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_integrate_thin -s 2 -c 1 -o gpurun_out/r02j_k3_thin_ml145 -f python profiles/r02_check.py ml145 > gpurun_out/r02j_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_layers -s 2 -c 1 -o gpurun_out/r02j_k0_ml145 -f python profiles/r02_check.py ml145 > gpurun_out/r02j_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sample_stream -s 3 -c 1 -o gpurun_out/r02j_k2 -f python bench.py --steps 1 --warmup 1 > gpurun_out/r02j_ncu5.log 2>&1
tail -3 gpurun_out/r02j_tests.log; cat gpurun_out/r02j_check.log; tail -3 gpurun_out/r02j_bench.err; cat gpurun_out/r02j_bench_reference.json | cut -c1-400
