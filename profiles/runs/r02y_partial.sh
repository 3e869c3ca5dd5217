#!/bin/bash
# round 2, run y: partial-level staging of the thin-layer kernel (footprints beyond the capacity)
set -x
mkdir -p gpurun_out
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 2>&1 | tail -3 | cut -c1-330
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['variants'].items(): print(k, v['ms_per_step'], v['ms_k0'], v['ms_k3'], v['tma_staged_passes'], v['max_abs_diff_vs_proj_form_m'])
P
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
