# launch list with DRAM / tensor metrics over bench steps (final state of this session)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 400 -c 220 --csv --log-file gpurun_out/launches_r1c_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "ncu metrics rc=$?"
python tools/profile_summary.py gpurun_out/launches_r1c_metrics.csv | head -30
