timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_driver.json 2> gpurun_out/r2_bench_driver.err; echo "bench exit=$?"; tail -3 gpurun_out/r2_bench_driver.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_driver.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'launches/fwd', d['launches_per_forward'])
for k in d['kernels']:
    if k['kernel'] in ('decode_nms3d', 'fps_sa1', 'proposal_fps_ballquery_group_mlp'): print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('default: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1))
PY
