#!/bin/bash
# One GPU call that decides the two unmeasured round-1 candidates:
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/measure_candidates.sh > gpurun_out/candidates.log 2>&1; tail -40 gpurun_out/candidates.log'
# 1. parity of both candidates (exact equality with the default paths), 2. van Hove kernel: default vs MDB_HIST_F32=1,
# 3. dense kernel: default staging vs MDB_DENSE_STAGE=warp.  Make a candidate the default only if its tests pass AND it wins.
set -u
cd "$(dirname "$0")/.."
echo "== parity of the candidates"
MDB_TEST_CANDIDATES=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k candidate 2>&1 | tail -5
echo "== van Hove, config-2 shape (S=100): default"
timeout 120 python scripts/kernel_rates.py 100000 10000 100 2>&1 | grep -E "^vh|fused"
echo "== van Hove: MDB_HIST_F32=1"
MDB_HIST_F32=1 timeout 120 python scripts/kernel_rates.py 100000 10000 100 2>&1 | grep -E "^vh|fused"
echo "== dense schedule (N=1e5, T=4000): default staging"
timeout 200 python scripts/dense_rates.py 100000 4000 2>&1 | grep tiled
echo "== dense schedule: MDB_DENSE_STAGE=warp"
MDB_DENSE_STAGE=warp timeout 200 python scripts/dense_rates.py 100000 4000 2>&1 | grep tiled
