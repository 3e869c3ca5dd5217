#!/bin/bash
# round 2, run p: span length of the polynomial geometry (RDR_K3_SPAN) x thick-tail absorb
set -x
mkdir -p gpurun_out
for sp in 12000 18000 24000 36000; do
for a in 0 128; do
RDR_K3_SPAN=$sp RDR_K3_THIN_ABSORB=$a timeout 300 python profiles/r02_check.py ml145 c2 2>&1 | tail -2 | cut -c1-330
done
done
