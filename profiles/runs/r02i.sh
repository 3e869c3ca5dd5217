set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 > gpurun_out/r02i_check.log 2>&1
for v in "RDR_K3_UNIFIED=1" "RDR_K3_UNIFIED=1 RDR_K3_THIN_MINB=3" "RDR_K3_THIN_MINB=3" "RDR_K3_UNIFIED=1 RDR_K3_THIN_TILE=0" "RDR_K3_UNIFIED=1 RDR_K3_QUAD=0"; do
  echo "== $v" >> gpurun_out/r02i_variants.log
  env $v timeout 200 python profiles/r02_check.py c2 ml145 >> gpurun_out/r02i_variants.log 2>&1
done
RDR_K3_UNIFIED=1 timeout 900 python -m pytest tests -m gpu -x -q -k "golden or fused or thin or forms or predicates or nan_nodes or live" > gpurun_out/r02i_tests_unified.log 2>&1
RDR_K3_UNIFIED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_integrate_thin -s 2 -c 1 -o gpurun_out/r02i_k3_unified_c2 -f python profiles/r02_check.py c2 > gpurun_out/r02i_ncu.log 2>&1
cat gpurun_out/r02i_check.log; tail -3 gpurun_out/r02i_tests_unified.log
