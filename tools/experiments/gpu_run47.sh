for n in 128 96 64; do echo "PAIR_MIN_N $n"; B2D_PAIR_MIN_N=$n python tools/diag.py time --batch 64 > gpurun_out/d_time_pair$n.log 2>&1; tail -2 gpurun_out/d_time_pair$n.log | head -1; done
B2D_PAIR_MIN_N=64 python -m pytest tests/test_gpu_parity.py -x -q -k "planned_op" 2>&1 | tail -2
