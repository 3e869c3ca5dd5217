#!/bin/bash
# round 2, run v: thin-layer kernel with the ray frame rebuilt per span (registers freed in the layer loop)
set -x
timeout 300 python profiles/r02_check.py c2 ml145 hrrr57 2>&1 | tail -3 | cut -c1-330
RDR_K3_THIN_MINB=5 timeout 300 python profiles/r02_check.py ml145 hrrr57 2>&1 | tail -2 | cut -c1-330
