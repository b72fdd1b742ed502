# NMS early-out: exactness tests + bench
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nms or postprocess or mosaic or full_batch or detect_host" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"
tail -c 200 gpurun_out/pytest_new.log
python bench.py --no-cpu-baseline > gpurun_out/bench_nms.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_nms.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e']['c_abi_one_call']['value'], d['roofline']['frac'], d['clocks'], d['roofline_aux']['postprocess']['ms_per_step'], d['roofline_aux']['preprocess']['ms_per_step'], d['config']['detections_last_step'])
PY
