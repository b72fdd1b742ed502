export B2D_LIB=tools/ubench/build/libb2det_trace.so
for e in 0 1 2 4 7; do echo "exp $e"; B2D_EXP=$e timeout 100 python tools/one_op.py --op 3 10 2 9 73 1 8 --reps 20 2>&1 | grep "^op"; done > gpurun_out/ablate3.log 2>&1
bash tools/gpu_trace.sh "3 10 2 9 73"
