mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_10.log
python tools/gpu_eye_timing.py 1024 > gpurun_out/r2/eye_timing_10.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_10.json 2> gpurun_out/r2/bench_10.err
tail -3 gpurun_out/r2/gpu_tests_10.log; cat gpurun_out/r2/eye_timing_10.log | tail -4; head -c 300 gpurun_out/r2/bench_10.json; echo; tail -5 gpurun_out/r2/bench_10.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2/bench_10.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step')}, d['e2e']['value'])
for k, v in (d.get('extra') or {}).items():
    print(k, v.get('value'), v.get('e2e', {}).get('value'), v.get('stage_ms_per_step'))
PY
