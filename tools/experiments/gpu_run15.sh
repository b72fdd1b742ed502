B2D_MT=4 timeout 200 python tools/diag.py tcops --batch 3 --imgsz 320 > gpurun_out/d_tcops.log 2>&1; echo "tcops mt rc=$?"
grep -c " ok " gpurun_out/d_tcops.log; grep -c BAD gpurun_out/d_tcops.log; grep -v " ok " gpurun_out/d_tcops.log | head -20 | cut -c1-220
B2D_MT=1 timeout 200 python tools/diag.py tcops --batch 3 --imgsz 320 > gpurun_out/d_tcops_mt1.log 2>&1; echo "tcops mt1 rc=$?"
grep -c " ok " gpurun_out/d_tcops_mt1.log; grep -c BAD gpurun_out/d_tcops_mt1.log; grep -v " ok " gpurun_out/d_tcops_mt1.log | head -20 | cut -c1-220
B2D_MT=4 timeout 200 python tools/diag.py tcops --arch yolov7 --batch 2 --imgsz 128 > gpurun_out/d_tcops_v7.log 2>&1; echo "tcops v7 rc=$?"
grep -c " ok " gpurun_out/d_tcops_v7.log; grep -c BAD gpurun_out/d_tcops_v7.log; grep -v " ok " gpurun_out/d_tcops_v7.log | head -20 | cut -c1-220
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time13.log 2>&1; tail -2 gpurun_out/d_time13.log | head -1
