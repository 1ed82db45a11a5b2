#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30
echo "== kernel rates, config-2 shape S=100"
timeout 300 python scripts/kernel_rates.py 100000 10000 100
echo "== host overhead S=20"
timeout 200 python scripts/host_overhead.py 20
echo "== bench N=1 (no config2)"
timeout 900 python bench.py --steps 20 --warmup 5 --no-config2 > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err; echo "rc=$?"; tail -3 gpurun_out/bench_n1_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_b.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.4g ms/step %.4f kernel %.4f frac %.3f api %.3f overhead %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['through_api']['frac'], r['through_api']['host_overhead_ms']))
print('sustained', r.get('sustained'))
print('other', d['other_kernels'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], 'parity', d['parity']['within_tolerance'], d['setup'])
PY
