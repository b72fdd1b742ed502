B2D_MT=4 timeout 200 python tools/diag.py tcops --batch 3 --imgsz 320 --only depthwise > gpurun_out/d_tcops.log 2>&1; echo "tcops mt rc=$?"
grep -c " ok " gpurun_out/d_tcops.log; grep -c BAD gpurun_out/d_tcops.log; grep -B1 -A3 BAD gpurun_out/d_tcops.log | cut -c1-200 | head -30
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time16.log 2>&1; tail -2 gpurun_out/d_time16.log | head -1
grep -E "depthwise" gpurun_out/d_time16.log | cut -c1-150
