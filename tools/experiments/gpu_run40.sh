export B2D_LIB=tools/ubench/build/libb2det_trace.so
for e in 0 8 16 32 64 96 88 120 2; do echo "exp $e"; B2D_EXP=$e timeout 100 python tools/one_op.py --op 2 9 3 10 18 --reps 20 2>&1 | grep "^op"; done > gpurun_out/ablate4.log 2>&1
