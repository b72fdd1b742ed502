OPS="2 9 21 3"
python tools/one_op.py --op $OPS --reps 1 > gpurun_out/oneop_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_tc' -s 82 -c 4 -o gpurun_out/prof_r1c_ops python tools/one_op.py --op $OPS --reps 1 > gpurun_out/oneop_ncu.log 2>&1
echo "ncu full rc=$?"
