#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="python scripts/overlap_rate.py 1000000 200 4"
$P > gpurun_out/plain_overlap.log 2>&1; tail -3 gpurun_out/plain_overlap.log
ncu --set full --clock-control none --import-source on -k regex:overlap_query_kernel -c 1 -f -o gpurun_out/r2_prof_overlap $P > gpurun_out/ncu_overlap.log 2>&1; echo "rc=$?"
