set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_tests.log
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02a_check.log 2>&1
for v in "RDR_K3_THIN_PF=0 RDR_K3_THIN_PFT=0" "RDR_K3_THIN_PF=0" "RDR_K3_THIN_PFT=0" "RDR_K3_THIN_PF=6" "RDR_K3_THIN_PFT=12" "RDR_K3_THIN_MINB=3" "RDR_K3_THIN_MINB=5" "RDR_K3_THIN_TILE=0" "RDR_K3_THIN_MIN=0" "RDR_K3_THIN_MIN=0 RDR_K3_CACHE=0 RDR_K3_TILE=0"; do
  echo "== $v" >> gpurun_out/r02a_variants.log
  env $v timeout 200 python profiles/r02_check.py ml145 >> gpurun_out/r02a_variants.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_integrate_thin -s 2 -c 1 -o gpurun_out/r02a_k3_thin_ml145 -f python profiles/r02_check.py ml145 > gpurun_out/r02a_ncu_thin.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_layers -s 2 -c 1 -o gpurun_out/r02a_k0_ml145 -f python profiles/r02_check.py ml145 > gpurun_out/r02a_ncu_k0.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_tests.log; cat gpurun_out/r02a_check.log
