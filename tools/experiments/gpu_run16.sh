for g in 148 111 74 37; do
echo "grid $g"; B2D_GRID=$g python tools/one_op.py --op 21 9 10 3 30 --reps 20 2>&1 | grep "^op"
done > gpurun_out/grid_exp.log 2>&1
python tools/one_op.py --op 21 --reps 1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'conv_tc' -s 82 -c 1 -o gpurun_out/prof_r1e_op21 python tools/one_op.py --op 21 --reps 1 > gpurun_out/oneop_ncu.log 2>&1
echo "ncu rc=$?"
