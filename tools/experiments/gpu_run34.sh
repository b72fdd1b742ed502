python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/launches_r1_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "ncu metrics rc=$?"
OPS="21 10 43 3"
python tools/one_op.py --op $OPS --reps 1 > gpurun_out/oneop_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_tc' -s 90 -c 4 -o gpurun_out/prof_r1_final_ops python tools/one_op.py --op $OPS --reps 1 > gpurun_out/oneop_ncu.log 2>&1
echo "ncu full rc=$?"
