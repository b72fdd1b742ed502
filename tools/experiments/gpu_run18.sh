B2D_MT=4 timeout 200 python tools/diag.py tcops --batch 3 --imgsz 320 > gpurun_out/d_tcops.log 2>&1; echo "tcops mt rc=$?"
grep -c " ok " gpurun_out/d_tcops.log; grep -c BAD gpurun_out/d_tcops.log; grep -B1 -A3 BAD gpurun_out/d_tcops.log | cut -c1-200 | head -30
timeout 200 python tools/diag.py tcops --batch 2 --imgsz 640 > gpurun_out/d_tcops640.log 2>&1; echo "tcops 640 rc=$?"
grep -c " ok " gpurun_out/d_tcops640.log; grep -c BAD gpurun_out/d_tcops640.log
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time15.log 2>&1; tail -2 gpurun_out/d_time15.log | head -1
grep -E "depthwise|model.2.cv|model.4.cv1|22.cv3.0.0.1" gpurun_out/d_time15.log | cut -c1-150
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
