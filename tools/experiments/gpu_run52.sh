python -m pytest tests/test_gpu_parity.py -x -q -k "forward_against or simple_detector or sharded or fused_head" > gpurun_out/pytest_ops.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ops.log
for g in 0 1; do echo "GRAPH $g"; B2D_GRAPH=$g python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_graph$g.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_graph$g.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['forward_ms'], d['clocks'])
PY
done
