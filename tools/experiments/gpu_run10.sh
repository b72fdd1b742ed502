for i in 1 2; do
B2D_LIB=tools/ubench/build/libb2det_head.so timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time8_head$i.log 2>&1; tail -2 gpurun_out/d_time8_head$i.log | head -1
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time8_new$i.log 2>&1; tail -2 gpurun_out/d_time8_new$i.log | head -1
done
