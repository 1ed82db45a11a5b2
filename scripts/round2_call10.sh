#!/bin/bash
# compute-sanitizer memcheck over the small cases of every kernel family
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/sanitizer_cases.py > gpurun_out/san_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitizer_cases.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/r2_sanitizer_memcheck.log
