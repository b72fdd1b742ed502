# b2d_detect_host with pinned result staging: parity + bench figure
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "detect_host" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"
tail -c 200 gpurun_out/pytest_new.log
python bench.py --no-cpu-baseline > gpurun_out/bench_cabi.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_cabi.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e']['c_abi_one_call'], d['roofline']['frac'], d['clocks'])
PY
