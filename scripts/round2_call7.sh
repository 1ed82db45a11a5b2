#!/bin/bash
# GPU call 7: full GPU test-suite, the default bench line, ncu launch list + --set full captures
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15
echo "== host overhead"
timeout 200 python scripts/host_overhead.py 20 2>&1 | tail -20
echo "== bench N=1 default"
( time timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ); echo "rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-config2 --no-sustained"
echo "== ncu launch list"
$B > gpurun_out/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"corr_|finish_|nojump_|images_pack" --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu --set full: ISF, config-4 shape (1M atoms, S=20, L=81)"
ncu --set full --clock-control none --import-source on -k regex:"corr_bulk_kernel<2, 4" -s 4 -c 1 -f -o gpurun_out/r2_prof_isf_config4 $B > gpurun_out/ncu_isf4.log 2>&1; echo "rc=$?"
P="python scripts/profile_one.py"
echo "== ncu --set full: fused van Hove + alpha_2 and van Hove, config-2 shape S=100"
$P fused 100000 10000 100 > gpurun_out/plain_fused.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"corr_bulk_kernel" -s 2 -c 1 -f -o gpurun_out/r2_prof_fused $P fused 100000 10000 100 > gpurun_out/ncu_fused.log 2>&1; echo "rc=$?"
$P vh 100000 10000 100 > gpurun_out/plain_vh.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"corr_bulk_kernel" -s 2 -c 1 -f -o gpurun_out/r2_prof_vh $P vh 100000 10000 100 > gpurun_out/ncu_vh.log 2>&1; echo "rc=$?"
$P msd 100000 10000 100 > gpurun_out/plain_msd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"corr_bulk_kernel" -s 2 -c 1 -f -o gpurun_out/r2_prof_msd $P msd 100000 10000 100 > gpurun_out/ncu_msd.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
