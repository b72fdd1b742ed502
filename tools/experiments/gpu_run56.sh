python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for p in bf16 fp16; do python bench.py --precision $p --no-cpu-baseline > gpurun_out/bench_$p.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$p.log').read().strip().splitlines()[-1])
print('$p', d['dtype'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
PY
done
