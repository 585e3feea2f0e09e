#!/bin/bash
# Round-2 profiling pass (recipe: /opt/skills/guides/B200_PROFILING.md): launch list with per-launch time, DRAM bytes and tensor-pipe
# activity of one training step, then one full capture per hot kernel (raw metrics + per-instruction source page; the binary reports are
# not kept).  Usage: tools/gpu_profile_r02.sh [tag]      -> tools/summarize_ncu.py <tag> turns gpurun_out/ into profiles/<tag>_*
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-configs"
timeout 200 $CMD > gpurun_out/plain_${TAG}.log 2>&1 || exit 1          # the program exits 0 without ncu first
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 1000 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit $?"
i=0
for KRE in 'k_wgrad_pair<\(int\)1,' 'k_wgrad_pair<\(int\)0, \(int\)160' 'k_wgrad_c<\(int\)0, \(int\)112' 'k_edge_step_c' 'k_edge_dgrad_c<\(bool\)0' \
           'k_lin<\(int\)150, \(int\)150, \(unsigned int\)6153>' 'k_lin<\(int\)150, \(int\)150, \(unsigned int\)640>' 'k_lin<\(int\)100, \(int\)200' 'k_gather_dsr_c'; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${KRE}" -s 3 -c 1 -o gpurun_out/prof_${TAG}_${i} -f $CMD > gpurun_out/ncu_full_${TAG}_${i}.log 2>&1
  echo "full capture ${KRE} exit $?"
  ncu -i gpurun_out/prof_${TAG}_${i}.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_${i}.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_${i}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/sass_${TAG}_${i}.csv.gz
  rm -f gpurun_out/prof_${TAG}_${i}.ncu-rep
  i=$((i+1))
done
ls -la gpurun_out/ | grep "_${TAG}_" | tail -24
