# detect_host parity + TTA golden CRC test + bench with roofline_aux / c_abi_one_call
python -m pytest tests/test_gpu_parity.py tests/test_gpu_tta.py -m gpu -x -q -k "detect_host or golden or tta" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"
tail -c 400 gpurun_out/pytest_new.log
python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_default.log
