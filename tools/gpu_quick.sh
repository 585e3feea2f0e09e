#!/bin/bash
# One GPU-box visit during development: parity tests (optional), short bench.  Usage: tools/gpu_quick.sh <tag> [tests: 0|1|pattern]
TAG=${1:-q}; TESTS=${2:-1}
mkdir -p gpurun_out
if [ "$TESTS" != "0" ]; then
  if [ "$TESTS" = "1" ]; then SEL="tests"; else SEL="tests -k $TESTS"; fi
  timeout 240 python -m pytest $SEL -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1
  echo "pytest exit $?"; tail -15 gpurun_out/pytest_${TAG}.log
fi
timeout 400 python bench.py --steps 10 --warmup 3 $BENCH_ARGS > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err
python tools/show_bench.py gpurun_out/bench_${TAG}.log 2>/dev/null | head -24
