for i in 1 2 3; do python tools/diag.py time --batch 64 > gpurun_out/d_time25_$i.log 2>&1; tail -2 gpurun_out/d_time25_$i.log | head -1; done
