#!/bin/bash
# bench.py at N GPUs the way the driver launches it; writes gpurun_out/r2_bench_n$N.json
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ); echo "rc=$?"
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.4g  ms/step %.4f  kernel %.4f  frac %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac']))
print('through_api', json.dumps(r['through_api']))
print('sustained', r.get('sustained'))
print('e2e', d['e2e'])
print('parity', d.get('parity'))
PY
