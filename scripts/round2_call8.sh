#!/bin/bash
# 8 GPUs: bench.py --gpus 8 (strong scaling of config 4) and --gpus 4
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
for n in 8 4; do
  echo "== bench --gpus $n"
  ( time timeout 900 python bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err ); echo "rc=$?"; tail -4 gpurun_out/r2_bench_n$n.err
done
