set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench2.log
OPS="1 2 3 4 7 9 10 11 18 21 22 29 36 72 73 74"
python tools/one_op.py --op $OPS --reps 1 > gpurun_out/oneop_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_tc|dwconv' -s 88 -c 16 -o gpurun_out/prof_r1b_ops python tools/one_op.py --op $OPS --reps 1 > gpurun_out/oneop_ncu.log 2>&1
echo "ncu full rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 110 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "ncu list rc=$?"
