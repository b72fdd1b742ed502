B2D_PAIR=2 timeout 100 python tools/diag.py tcops --batch 8 --imgsz 640 --only pair > gpurun_out/d_tcops_pair.log 2>&1; echo "tcops pair rc=$?"
grep -c " ok " gpurun_out/d_tcops_pair.log; grep -c BAD gpurun_out/d_tcops_pair.log
B2D_PAIR=2 B2D_MT=4 timeout 200 python tools/diag.py tcops --arch yolov7 --batch 2 --imgsz 256 > gpurun_out/d_tcops_v7.log 2>&1; echo "tcops v7 rc=$?"
grep -c " ok " gpurun_out/d_tcops_v7.log; grep -c BAD gpurun_out/d_tcops_v7.log; grep -c pair gpurun_out/d_tcops_v7.log
B2D_PAIR=0 timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time19a.log 2>&1; tail -2 gpurun_out/d_time19a.log | head -1
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time19b.log 2>&1; tail -2 gpurun_out/d_time19b.log | head -1
grep pair gpurun_out/d_time19b.log | cut -c1-100 | head -5
