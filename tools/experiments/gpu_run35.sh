# full gate: gpu parity tests, smoke, default bench, reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
tail -c 600 gpurun_out/pytest_gpu.log
nproc
