#!/bin/bash
# round 2, run o: thick tail of the 145-node table absorbed by the thin-layer kernel (RDR_K3_THIN_ABSORB = most extra samples)
set -x
mkdir -p gpurun_out
for a in 0 64 128; do
RDR_K3_THIN_ABSORB=$a timeout 300 python profiles/r02_check.py ml145 2>&1 | tail -1 | cut -c1-420
done
RDR_K3_THIN_ABSORB=128 RDR_K3_THIN_MINB=3 timeout 300 python profiles/r02_check.py ml145 2>&1 | tail -1 | cut -c1-420
