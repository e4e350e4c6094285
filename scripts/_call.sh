for IF in 16 20 24 32; do
for ST in 20 600; do
timeout 300 python bench.py --steps $ST --warmup 5 --no-cpu-baseline --inflight $IF > gpurun_out/sweep.json 2> gpurun_out/sweep.err || tail -2 gpurun_out/sweep.err
python - <<PY
import json
d = json.loads(open('gpurun_out/sweep.json').read().strip().splitlines()[-1])
print('inflight $IF steps $ST: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'lat1', round(d['latency_ms_inflight1'],3))
PY
done
done
