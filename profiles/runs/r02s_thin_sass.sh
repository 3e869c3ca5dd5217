#!/bin/bash
# round 2, run s: per-instruction view (ncu source page) of the thin-layer kernel and of K0 on the 145-node table
set -x
mkdir -p gpurun_out
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/tmp_$3 -f ${@:4} > gpurun_out/r02s_ncu_$3.log 2>&1
  python profiles/ncu_keys.py gpurun_out/tmp_$3.ncu-rep > gpurun_out/r02s_$3_ncu.txt 2>&1
  ncu -i gpurun_out/tmp_$3.ncu-rep --page source --csv > gpurun_out/r02s_$3_sass.csv 2>/dev/null
  python profiles/sass_summary.py gpurun_out/r02s_$3_sass.csv >> gpurun_out/r02s_$3_ncu.txt 2>&1
  gzip -f gpurun_out/r02s_$3_sass.csv; rm -f gpurun_out/tmp_$3.ncu-rep; }
cap k_ray_integrate_thin 2 k3_thin_ml145 python profiles/r02_check.py ml145
cap k_ray_layers 2 k0_ml145 python profiles/r02_check.py ml145
cap k_ray_integrate_poly 2 k3_poly_c2 python profiles/r02_check.py c2
du -sh gpurun_out
