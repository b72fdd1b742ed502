OPS="2 3 10 21 9 1"
python tools/one_op.py --op $OPS --reps 1 > gpurun_out/oneop_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_tc' -s 82 -c 6 -o gpurun_out/prof_r1d_ops python tools/one_op.py --op $OPS --reps 1 > gpurun_out/oneop_ncu.log 2>&1
echo "ncu full rc=$?"
