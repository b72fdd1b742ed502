export B2D_LIB=tools/ubench/build/libb2det_trace.so
for e in 0 8 16 32 64 120 2; do echo "exp $e"; B2D_EXP=$e timeout 100 python tools/one_op.py --op 0 3 4 2 9 10 73 --reps 20 2>&1 | grep "^op" | tr '\n' ' '; echo; done > gpurun_out/ablate5.log 2>&1
