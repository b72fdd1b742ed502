timeout 200 python tools/diag.py tcops --batch 3 --imgsz 320 > gpurun_out/d_tcops.log 2>&1; echo "tcops rc=$?"
grep -c " ok " gpurun_out/d_tcops.log; grep -c BAD gpurun_out/d_tcops.log
for v in mufu2 elect elect2; do
B2D_LIB=tools/ubench/build/libb2det_$v.so timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time10_$v.log 2>&1; echo $v; tail -2 gpurun_out/d_time10_$v.log | head -1
done
