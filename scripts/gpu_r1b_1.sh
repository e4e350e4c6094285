#!/bin/bash
# round 1 (session 2), call 1: pruned FPS parity + timing + bench
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 900 python -m pytest tests/test_gpu_fps_pruned.py -x -q 2>&1 | tail -15
timeout 600 python scripts/gpu_probe_fps.py 2>&1 | tee gpurun_out/probe_fps.txt
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_pruned.json 2> gpurun_out/bench_pruned.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_pruned.err
python - <<'PY'
import json
for f in ('bench_pruned',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'clocks', d['clocks'])
        for k in d['kernels']: print('   ', k['kernel'], k['ms'], round(k['frac'],3))
    except Exception as e: print(f, 'failed', e)
PY
for inf in 2 4 16; do
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --inflight $inf 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight $inf: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
done
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --fps-variant 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old fps (256x4): value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
