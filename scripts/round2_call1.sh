#!/bin/bash
# GPU call 1 of round 2: baseline data for the design decisions (see DESIGN.md section 9 of round 1).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== shared-memory histogram update micro-benchmark"
timeout 120 scripts/_bin/ubench_atoms
echo "== candidates of round 1"
timeout 900 bash scripts/measure_candidates.sh
echo "== small launches: config-4 per-GPU shape at 8 GPUs (125k atoms)"
timeout 300 python scripts/small_launch_sweep.py 125000 4000
echo "== small launches: 250k atoms (4 GPUs)"
timeout 300 python scripts/small_launch_sweep.py 250000 4000 2>&1 | grep -E "default|ABLOCKS=(4|8|16)\b"
echo "== host overhead"
timeout 120 python scripts/host_overhead.py 20
