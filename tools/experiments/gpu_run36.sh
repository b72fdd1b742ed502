python tools/mosaic_bench.py --size 10000 --repeat 2 > gpurun_out/mosaic_10k_n1.log 2>&1; echo "m10k rc=$?"; tail -2 gpurun_out/mosaic_10k_n1.log
python tools/mosaic_bench.py --size 40000 --repeat 1 > gpurun_out/mosaic_40k_n1.log 2>&1; echo "m40k rc=$?"; tail -2 gpurun_out/mosaic_40k_n1.log
