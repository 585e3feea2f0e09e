#!/bin/bash
# full ncu capture of one launch of each named kernel: tools/gpu_ncu.sh <tag> <kernel-regex> [<kernel-regex> ...]
TAG=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3"
i=0
for KRE in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 6 -c 1 -o gpurun_out/prof_${TAG}_${i} -f $CMD > gpurun_out/ncu_full_${TAG}_${i}.log 2>&1
  echo "full capture ${KRE} exit $?"
  i=$((i+1))
done
ls -la gpurun_out/prof_${TAG}_* | tail
