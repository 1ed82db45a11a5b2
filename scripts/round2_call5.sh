#!/bin/bash
# 2 GPUs: the peer-memory communicator (tests/test_gpu_multi.py) and bench.py --gpus 2
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | head -12
echo "== 2-GPU test"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --tb=long 2>&1 | tail -60
echo "== bench --gpus 2"
( time timeout 1200 python bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ); echo "rc=$?"; tail -15 gpurun_out/bench_n2.err
