#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guard_bands.py -m gpu -q --tb=short -k "dense" 2>&1 | tail -3
timeout 400 python scripts/dense_rates.py 100000 4000 2>&1 | grep "tiled v2"
echo "== config 5 at full size"
timeout 1500 python scripts/config5_full.py 200000 50000 gpurun_out/r2_config5_full.json 2>&1 | tail -8
