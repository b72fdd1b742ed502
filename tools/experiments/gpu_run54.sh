python -m pytest tests/test_gpu_parity.py -x -q -k "mosaic or sharded" > gpurun_out/pytest_ops.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ops.log
python tools/mosaic_bench.py --size 40000 --repeat 2 > gpurun_out/mosaic_40k_n1.log 2>&1; echo "m40k rc=$?"; grep '^{' gpurun_out/mosaic_40k_n1.log | cut -c1-420
