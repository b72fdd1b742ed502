python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "ncu list rc=$?"
