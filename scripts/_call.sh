timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit=$?"; tail -2 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('default: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'lat1', d['latency_ms_inflight1'])
PY
