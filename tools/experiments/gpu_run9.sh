timeout 200 python tools/diag.py tcops --batch 3 --imgsz 160 > gpurun_out/d_tcops.log 2>&1; echo "tcops rc=$?"
grep -c " ok " gpurun_out/d_tcops.log; grep -c BAD gpurun_out/d_tcops.log
timeout 200 python tools/diag.py tcops --arch yolov7 --batch 2 --imgsz 128 > gpurun_out/d_tcops_v7.log 2>&1; echo "tcops v7 rc=$?"
grep -c " ok " gpurun_out/d_tcops_v7.log; grep -c BAD gpurun_out/d_tcops_v7.log
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time7.log 2>&1; echo "time rc=$?"
tail -3 gpurun_out/d_time7.log
