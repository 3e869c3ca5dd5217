set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_tests.log
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02h_check.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
timeout 600 python profiles/run_configs.py > gpurun_out/r02h_configs.json 2> gpurun_out/r02h_configs.err
tail -5 gpurun_out/r02h_tests.log; cat gpurun_out/r02h_check.log; tail -3 gpurun_out/r02h_bench.err; tail -5 gpurun_out/r02h_configs.err
