#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
free -g | head -2; nproc
echo "== config 5 at full size"
timeout 1500 python scripts/config5_full.py 200000 50000 gpurun_out/r2_config5_full.json 2>&1 | tail -8
echo "== ncu: dense v2 kernels (N=1e5, T=1200: S=L=600)"
P="python scripts/profile_dense.py 100000 1200"
$P > gpurun_out/plain_dense.log 2>&1; tail -3 gpurun_out/plain_dense.log
ncu --set full --clock-control none --import-source on -k regex:dense_corr_kernel_v2 -c 2 -f -o gpurun_out/r2_prof_dense_v2 $P > gpurun_out/ncu_dense.log 2>&1; echo "rc=$?"
