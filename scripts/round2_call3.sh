#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== debug shard"
timeout 300 python scripts/debug_shard.py
echo "== gpu tests (all, no -x)"
timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150
