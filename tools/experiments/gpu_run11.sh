for v in head notrace mufu2; do
B2D_LIB=tools/ubench/build/libb2det_$v.so timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time9_$v.log 2>&1; echo $v; tail -2 gpurun_out/d_time9_$v.log | head -1
done
