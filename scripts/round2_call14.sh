#!/bin/bash
# the N=1 line for the record + the ncu launch list of the same command (kernel share of a step)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ); echo "bench rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-config2 --no-sustained"
$B > gpurun_out/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"corr_|finish_|nojump_|images_pack" --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
grep -c . gpurun_out/r2_launches_bench.csv
