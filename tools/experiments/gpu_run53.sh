python -m pytest tests/test_gpu_parity.py -x -q -k "forward_against or simple_detector or sharded or fused_head or onnx" > gpurun_out/pytest_ops.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ops.log
for ns in 1 2 3 4; do echo "STREAMS $ns"; B2D_STREAMS=$ns python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_streams$ns.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_streams$ns.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['forward_ms'], d['clocks'])
PY
done
