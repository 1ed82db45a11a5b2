#!/bin/bash
# final N=1 record: smoke, both bench arms, the ncu launch list of the bench command
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 1200 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ); echo "ref rc=$?"
bash scripts/round2_call14.sh
