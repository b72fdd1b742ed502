for ps in 0 1 2 3; do echo "promo_slice $ps"; B2D_PROMO_SLICE=$ps python tools/one_op.py --op 3 4 10 11 51 19 30 32 21 23 --reps 20 2>&1 | grep "^op" | tr '\n' ' '; echo; done > gpurun_out/promo.log 2>&1
for pd in 0 2; do echo "promo_dense $pd"; B2D_PROMO_DENSE=$pd python tools/one_op.py --op 1 2 7 9 18 29 52 --reps 20 2>&1 | grep "^op" | tr '\n' ' '; echo; done >> gpurun_out/promo.log 2>&1
python tools/diag.py time --batch 64 > gpurun_out/d_time23.log 2>&1; tail -3 gpurun_out/d_time23.log
