python tools/diag.py time --arch yolov7 --batch 128 > gpurun_out/d_time_v7_b128.log 2>&1; echo "v7 rc=$?"; tail -3 gpurun_out/d_time_v7_b128.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
