#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40
echo "== small bench (flow check)"
timeout 600 python bench.py --atoms 250000 --frames 4000 --steps 5 --warmup 2 --cpu-atoms 2048 --e2e-steps 2 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?"; tail -5 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json | cut -c1-3000
echo "== full bench N=1"
( time timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err ); echo "rc=$?"; tail -5 gpurun_out/bench_full_n1.err
