mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_h1.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_h1.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_h1.log 2> gpurun_out/bench_h1.err; echo "bench exit $?"
python tools/show_bench.py gpurun_out/bench_h1.log | grep -v "^  k_[a-z_:0-9]* .* 0\.[0-9]%$"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_h1.log 2>&1; echo "ref exit $?"; tail -c 600 gpurun_out/bench_ref_h1.log
CMD="python bench.py --steps 1 --warmup 3 --no-configs"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_edge_step_c" -s 3 -c 1 -o gpurun_out/prof_r02_3 -f $CMD > gpurun_out/ncu_full_r02_3.log 2>&1
echo "capture exit $?"
ncu -i gpurun_out/prof_r02_3.ncu-rep --page raw --csv > gpurun_out/raw_r02_3.csv 2>/dev/null
ncu -i gpurun_out/prof_r02_3.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/sass_r02_3.csv.gz
rm -f gpurun_out/prof_r02_3.ncu-rep
