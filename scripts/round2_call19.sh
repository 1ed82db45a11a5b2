#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6
( time timeout 1200 python bench.py --no-cpu > gpurun_out/r2_bench_n1_fast.json 2> gpurun_out/r2_bench_n1_fast.err ); echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n1_fast.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.4g ms %.4f kernel %.4f (min %.4f med %.4f) frac %.3f overhead %.4f'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_median'],r['frac'],r['through_api']['host_overhead_ms']))
print('sustained', r['sustained']['frac'], r['sustained']['value'], r['sustained']['clocks'])
print('other', json.dumps(d['other_kernels']))
c2=d['config2']
print('c2', c2['value'], c2['roofline']['frac'], json.dumps(c2['other_kernels']))
print('parity', d['parity'])
PY
