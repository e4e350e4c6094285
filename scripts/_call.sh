timeout 200 python scripts/gpu_nms_probe.py 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_index_ops.py tests/test_gpu_forward.py tests/test_gpu_engine.py tests/test_gpu_concurrent.py -x -q 2>&1 | tail -n 4
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit=$?"; tail -2 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('default: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'lat1', d['latency_ms_inflight1'], d['configs1_single_sa_layer'])
for k in d['kernels']:
    print('   %-36s %8.4f ms  frac %.3f  %s' % (k['kernel'], k['ms'], k['frac'], k.get('frac_executed', '')))
PY
