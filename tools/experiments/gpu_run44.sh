python -m pytest tests/test_gpu_parity.py -x -q -k "planned_op or forward_against" > gpurun_out/pytest_ops.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ops.log
python tools/diag.py time --batch 64 > gpurun_out/d_time24.log 2>&1; tail -2 gpurun_out/d_time24.log | head -1
