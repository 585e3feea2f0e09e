#!/bin/bash
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 3 --no-configs"
i=0
for KRE in 'k_wgrad_c<\(int\)1,' 'k_wgrad_c<\(int\)0, \(int\)160' 'k_edge_dgrad_c<\(bool\)0' 'k_lin<\(int\)150, \(int\)150, \(unsigned int\)6153>' 'k_lin<\(int\)150, \(int\)150, \(unsigned int\)640>' 'k_lin<\(int\)100, \(int\)200'; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${KRE}" -s 3 -c 1 -o gpurun_out/prof_${TAG}b_${i} -f $CMD > gpurun_out/ncu_full_${TAG}b_${i}.log 2>&1
  echo "full capture ${KRE} exit $?"
  ncu -i gpurun_out/prof_${TAG}b_${i}.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}b_${i}.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}b_${i}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/sass_${TAG}b_${i}.csv.gz
  rm -f gpurun_out/prof_${TAG}b_${i}.ncu-rep
  i=$((i+1))
done
ls -la gpurun_out/ | grep ${TAG}b | tail -14
