#!/bin/bash
# ncu --set full of the final ISF kernel on the bench's shape (config 4) -> roofline.traffic
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-config2 --no-sustained"
$B > gpurun_out/plain_isf4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:corr_bulk_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_isf_config4_final $B > gpurun_out/ncu_isf4.log 2>&1; echo "rc=$?"
