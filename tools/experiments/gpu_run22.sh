timeout 200 python tools/diag.py tcops --batch 2 --imgsz 640 > gpurun_out/d_tcops640.log 2>&1; echo "tcops 640 rc=$?"
grep -c " ok " gpurun_out/d_tcops640.log; grep -c BAD gpurun_out/d_tcops640.log
B2D_PDL=0 timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time18a.log 2>&1; tail -2 gpurun_out/d_time18a.log | head -1
timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time18b.log 2>&1; tail -2 gpurun_out/d_time18b.log | head -1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
