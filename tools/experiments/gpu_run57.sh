# full gate after the test-time-augmentation row: gpu parity tests (all files), smoke, default bench, reference arm, TTA bench
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
python tools/tta_bench.py > gpurun_out/tta_bench3.json 2> gpurun_out/tta_bench.err; echo "tta bench rc=$?"
tail -c 300 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; tail -1 gpurun_out/bench_default.log; tail -1 gpurun_out/bench_ref.log; cat gpurun_out/tta_bench3.json
