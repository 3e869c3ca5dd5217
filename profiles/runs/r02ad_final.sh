# round 2, run ad: the final tree after the upper clamp (ABI 3).  NGPU=1: tests, K2 traffic, bench, ncu of K0; NGPU>1: dist_check + bench lines
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NGPU:-1}
if [ "$N" = "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02ad_gpu_tests.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_sample_stream --csv --log-file gpurun_out/r02ad_k2_dram.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r02ad_k2_dram_bench.log 2>&1
python profiles/ncu_traffic.py gpurun_out/r02ad_k2_dram.csv 192000000 > gpurun_out/k2_traffic.json 2> gpurun_out/r02ad_k2_traffic.err
cp gpurun_out/k2_traffic.json profiles/k2_traffic.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ad_bench.json 2> gpurun_out/r02ad_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ad_bench_reference.json 2> gpurun_out/r02ad_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ad_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02ad_launches_bench.log 2>&1
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/tmp_$3 -f ${@:4} > gpurun_out/r02ad_ncu_$3.log 2>&1; bash profiles/summarize_ncu.sh gpurun_out/tmp_$3.ncu-rep gpurun_out/r02ad_$3_ncu.txt; }
cap k_ray_layers 2 k0_c2 python profiles/r02_check.py c2
cap k_ray_layers 2 k0_ml145 python profiles/r02_check.py ml145
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "fused or staged or knife or all_nan or slant or clamp" > gpurun_out/r02ad_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02ad_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "staged or fused or clamp" > gpurun_out/r02ad_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02ad_racecheck.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ad_smoke.txt 2>&1
cat gpurun_out/r02ad_gpu_tests.txt; tail -3 gpurun_out/r02ad_memcheck.log; tail -3 gpurun_out/r02ad_racecheck.log; cat gpurun_out/r02ad_smoke.txt; cut -c1-300 gpurun_out/r02ad_bench.json
else
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 ${@:3} > gpurun_out/r02ad_$2_n$N.json 2> gpurun_out/r02ad_$2_n$N.err; }
run 29511 dist_check profiles/dist_check.py
run 29512 bench bench.py --gpus $N --steps 10 --warmup 3
run 29513 bench_c5 bench.py --gpus $N --steps 5 --warmup 3 --config c5
cut -c1-1200 gpurun_out/r02ad_dist_check_n$N.json; tail -2 gpurun_out/r02ad_dist_check_n$N.err; cut -c1-200 gpurun_out/r02ad_bench_n$N.json
fi
