B2D_PAIR=2 timeout 100 python tools/diag.py tcops --batch 8 --imgsz 640 --only pair > gpurun_out/d_tcops_pair.log 2>&1; echo "tcops pair rc=$?"
grep -c " ok " gpurun_out/d_tcops_pair.log; grep -c BAD gpurun_out/d_tcops_pair.log; grep -A6 BAD gpurun_out/d_tcops_pair.log | cut -c1-220 | head -24
bash tools/gpu_trace.sh "21"
