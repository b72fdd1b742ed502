python bench.py --steps 20 --warmup 5 > gpurun_out/bench3.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench3.log
timeout 300 python tools/diag.py time --arch yolov7 --batch 64 > gpurun_out/d_time_v7.log 2>&1; tail -2 gpurun_out/d_time_v7.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
