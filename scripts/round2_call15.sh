#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="python scripts/profile_dense.py 100000 1200"
$P > gpurun_out/plain_dense.log 2>&1; tail -3 gpurun_out/plain_dense.log
ncu --set full --clock-control none --import-source on -k regex:dense_corr_kernel_v2 -c 2 -f -o gpurun_out/r2_prof_dense_v2b $P > gpurun_out/ncu_dense.log 2>&1; echo "rc=$?"
