B2D_PAIR=2 timeout 60 python tools/diag.py tcops --batch 3 --imgsz 320 --only pair > gpurun_out/d_tcops_pair.log 2>&1; echo "tcops pair rc=$?"
grep -c " ok " gpurun_out/d_tcops_pair.log; grep -c BAD gpurun_out/d_tcops_pair.log; grep -B1 -A6 BAD gpurun_out/d_tcops_pair.log | cut -c1-220 | head -40; tail -3 gpurun_out/d_tcops_pair.log | cut -c1-200
timeout 60 python tools/one_op.py --op 21 22 44 --reps 20 2>&1 | tail -6
