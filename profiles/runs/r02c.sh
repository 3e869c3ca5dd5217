set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_tests.log
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02c_check.log 2>&1
for v in "RDR_K3_THIN_STAGE=0" "RDR_K3_THIN_MINB=3" "RDR_K3_THIN_MINB=5" "RDR_K3_THIN_TILE=0" "RDR_K3_THIN_TILE=2" "RDR_K3_THIN_PFT=0" "RDR_K3_THIN_PFT=12" "RDR_K3_SPAN=8000" "RDR_K3_SPAN=16000"; do
  echo "== $v" >> gpurun_out/r02c_variants.log
  env $v timeout 200 python profiles/r02_check.py ml145 >> gpurun_out/r02c_variants.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_integrate_thin -s 2 -c 1 -o gpurun_out/r02c_k3_thin_ml145 -f python profiles/r02_check.py ml145 > gpurun_out/r02c_ncu_thin.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_integrate_poly -s 4 -c 1 -o gpurun_out/r02c_k3_poly_ml145 -f python profiles/r02_check.py ml145 > gpurun_out/r02c_ncu_poly.log 2>&1
tail -3 gpurun_out/r02c_tests.log; cat gpurun_out/r02c_check.log
