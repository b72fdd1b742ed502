export B2D_LIB=tools/ubench/build/libb2det_trace.so
for e in 0 1 2 3 4 5 6 7; do echo "exp $e"; B2D_EXP=$e timeout 60 python tools/one_op.py --op 21 10 3 --reps 20 2>&1 | grep "^op"; done > gpurun_out/ablate.log 2>&1
