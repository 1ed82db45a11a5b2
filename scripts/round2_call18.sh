#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8
bash scripts/round2_call14.sh
