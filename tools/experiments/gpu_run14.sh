timeout 200 python tools/diag.py tcops --batch 3 --imgsz 320 > gpurun_out/d_tcops.log 2>&1; echo "tcops rc=$?"
grep -c " ok " gpurun_out/d_tcops.log; grep -c BAD gpurun_out/d_tcops.log; head -3 gpurun_out/d_tcops.log | cut -c1-200
timeout 200 python tools/diag.py tcops --arch yolov7 --batch 2 --imgsz 128 > gpurun_out/d_tcops_v7.log 2>&1; echo "tcops v7 rc=$?"
grep -c " ok " gpurun_out/d_tcops_v7.log; grep -c BAD gpurun_out/d_tcops_v7.log; head -3 gpurun_out/d_tcops_v7.log | cut -c1-200
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time12.log 2>&1; tail -2 gpurun_out/d_time12.log | head -1; head -3 gpurun_out/d_time12.log | cut -c1-150
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
