CMD="python bench.py --steps 1 --warmup 3 --no-configs"
timeout 200 $CMD > gpurun_out/plain_r02.log 2>&1 || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches_r02.log 2>&1
echo "launch list exit $?"
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_h2.log 2> gpurun_out/bench_h2.err; echo "bench exit $?"
python tools/show_bench.py gpurun_out/bench_h2.log | grep -v "^  k_"
