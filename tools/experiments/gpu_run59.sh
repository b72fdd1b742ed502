# coalesced identity preprocess: parity tests that cover it + bench
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "preprocess or session_input or full_batch or simple_detector or gpu_handler" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"
tail -c 300 gpurun_out/pytest_new.log
python bench.py --no-cpu-baseline > gpurun_out/bench_prep.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_prep.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e']['c_abi_one_call']['value'], d['roofline']['frac'], d['clocks'], d['roofline_aux']['preprocess'])
PY
