N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tools/mosaic_bench.py --size 40000 --repeat 2 > gpurun_out/mosaic_40k_n$N.log 2>&1; echo "m40k rc=$?"; grep '^{' gpurun_out/mosaic_40k_n$N.log
$TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_n$N.log | cut -c1-330
