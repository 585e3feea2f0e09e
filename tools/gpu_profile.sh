#!/bin/bash
# ncu passes (recipe: /opt/skills/guides/B200_PROFILING.md).  Usage: tools/gpu_profile.sh <tag> [kernel-regex]
TAG=${1:-r01}
KRE=${2:-k_edge_step_bwd}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 5 -c 2 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/ | tail -12
