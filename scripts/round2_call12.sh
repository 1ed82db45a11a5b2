#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25
echo "== bench N=1"
( time timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ); echo "rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
echo "== bench reference arm"
( time timeout 1200 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ); echo "rc=$?"; tail -3 gpurun_out/r2_bench_ref.err; cut -c1-600 gpurun_out/r2_bench_ref.json
