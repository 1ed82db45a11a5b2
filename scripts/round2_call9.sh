#!/bin/bash
# ncu --set full of the ISF kernel on the config-4 shape (the bench's dominant kernel) -> roofline.traffic
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-config2 --no-sustained"
$B > gpurun_out/plain_isf4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:corr_bulk_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_isf_config4 $B > gpurun_out/ncu_isf4.log 2>&1; echo "rc=$?"
P="python scripts/profile_one.py"
$P isf 100000 10000 100 > gpurun_out/plain_isf2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:corr_bulk_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_isf_config2 $P isf 100000 10000 100 > gpurun_out/ncu_isf2.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r2_prof_isf*.ncu-rep
echo "== sanitizer cases, plain (must pass before any sanitizer run)"
timeout 600 python scripts/sanitizer_cases.py 2>&1 | tail -5
