#!/bin/bash
# Round-2 profiling pass (recipe: /opt/skills/guides/B200_PROFILING.md): launch list with per-launch time and DRAM bytes of one
# training step, then one full capture per hot kernel.  Usage: tools/gpu_profile_r02.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-configs"
timeout 200 $CMD > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 1000 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit $?"
i=0
for KRE in "k_wgrad_c<1" "k_wgrad_c<0, 160" k_edge_step_c "k_edge_dgrad_c<false" "k_lin<150, 150, 6153" "k_lin<150, 150, 640" "k_lin<100, 200"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:${KRE}" -s 3 -c 1 -o gpurun_out/prof_${TAG}_${i} -f $CMD > gpurun_out/ncu_full_${TAG}_${i}.log 2>&1
  echo "full capture ${KRE} exit $?"
  i=$((i+1))
done
ls -la gpurun_out/ | grep ${TAG} | tail -12
