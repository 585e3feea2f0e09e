#!/bin/bash
# extra full captures: tools/gpu_profile2.sh <tag> <kernel-regex> [skip]
TAG=$1; KRE=$2; SKIP=${3:-2}
CMD="python bench.py --steps 1 --warmup 3"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP} -c 1 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture ${TAG} exit $?"
