#!/bin/bash
# round 2, run aa (experiment, not kept): the quartic chord term of the layer quadrature compiled out of k_ray_integrate_poly
# (e4w = e4h = 0): C2 K3 2.54 -> 2.42 ms, max |poly - PROJ-form| 1.6e-10 -> 2.2e-10 m on C2 but 1.77e-9 m in
# tests/test_gpu_parity.py::test_k3_integrator_forms_agree (62 deg incidence, 500 m segments) -> the term stays.
# As a run-time flag (quad & 2) the kernel spills: 3.10 ms with the term, 2.93 ms without.
set -x
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 2>&1 | tail -3 | cut -c1-400
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "^(FAILED|E  +Assert|E  +assert|[0-9]+ (passed|failed))" | cut -c1-300
