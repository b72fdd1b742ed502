B2D_MT=4 timeout 200 python tools/diag.py tcops --batch 3 --imgsz 320 > gpurun_out/d_tcops.log 2>&1; echo "tcops mt rc=$?"
grep -c " ok " gpurun_out/d_tcops.log; grep -c BAD gpurun_out/d_tcops.log; grep -B1 -A3 BAD gpurun_out/d_tcops.log | cut -c1-200 | head -30
timeout 200 python tools/diag.py tcops --batch 2 --imgsz 640 > gpurun_out/d_tcops640.log 2>&1; echo "tcops 640 rc=$?"
grep -c " ok " gpurun_out/d_tcops640.log; grep -c BAD gpurun_out/d_tcops640.log
B2D_MT=4 timeout 200 python tools/diag.py tcops --arch yolov7 --batch 2 --imgsz 128 > gpurun_out/d_tcops_v7.log 2>&1; echo "tcops v7 rc=$?"
grep -c " ok " gpurun_out/d_tcops_v7.log; grep -c BAD gpurun_out/d_tcops_v7.log
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time17.log 2>&1; tail -2 gpurun_out/d_time17.log | head -1
