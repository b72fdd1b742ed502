export B2D_LIB=tools/ubench/build/libb2det_trace.so
for e in 0 8 2 32 120; do B2D_EXP=$e timeout 120 python tools/diag.py time --batch 64 > gpurun_out/d_time_exp$e.log 2>&1; echo "exp $e: $(tail -2 gpurun_out/d_time_exp$e.log | head -1)"; done
