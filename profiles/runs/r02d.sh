set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_tests.log
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02d_check.log 2>&1
for v in "RDR_K3_THIN_MINB=3" "RDR_K3_THIN_TILE=0" "RDR_K3_THIN_STAGE=0"; do
  echo "== $v" >> gpurun_out/r02d_variants.log
  env $v timeout 200 python profiles/r02_check.py ml145 >> gpurun_out/r02d_variants.log 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_integrate_thin -s 2 -c 1 -o gpurun_out/r02d_k3_thin_ml145 -f python profiles/r02_check.py ml145 > gpurun_out/r02d_ncu_thin.log 2>&1
tail -5 gpurun_out/r02d_tests.log; cat gpurun_out/r02d_check.log; tail -3 gpurun_out/r02d_bench.err
