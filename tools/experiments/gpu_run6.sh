export B2D_TRACE=1 B2D_TRACE_DUMP=1
for op in 2 9 3 10 21 7; do
python tools/one_op.py --op $op --reps 1 2>&1 | tail -30
done > gpurun_out/trace.log 2>&1
