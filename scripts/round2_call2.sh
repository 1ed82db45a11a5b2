#!/bin/bash
# GPU call 2 of round 2: the rewritten kernels (packed image counts, plan API, van Hove estimate) -- tests and rates.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke"
timeout 300 python __graft_entry__.py smoke
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
echo "== kernel rates, config-2 shape S=100"
timeout 300 python scripts/kernel_rates.py 100000 10000 100
echo "== host overhead S=20"
timeout 120 python scripts/host_overhead.py 20
