for mn in 128 48; do
echo "pair min n $mn"; B2D_PAIR_MIN_N=$mn timeout 300 python tools/diag.py time --batch 64 > gpurun_out/d_time20_$mn.log 2>&1; tail -2 gpurun_out/d_time20_$mn.log | head -1
done
B2D_PAIR_MIN_N=48 B2D_PAIR=2 timeout 100 python tools/diag.py tcops --batch 8 --imgsz 640 --only pair > gpurun_out/d_tcops_pair.log 2>&1; echo "tcops pair rc=$?"
grep -c " ok " gpurun_out/d_tcops_pair.log; grep -c BAD gpurun_out/d_tcops_pair.log
