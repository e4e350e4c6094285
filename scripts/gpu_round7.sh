#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_fps_prof.py 2>&1 | tee gpurun_out/fps_prof.log
for cfg in "256 8 4" "256 8 8" "512 4 8" "256 4 8"; do
set -- $cfg
timeout 600 python bench.py --steps 200 --warmup 10 --fps-threads $1 --fps-cluster $2 --inflight $3 --no-cpu-baseline > gpurun_out/bench_$1_$2_$3.json 2> gpurun_out/bench_$1_$2_$3.err
python - "$1" "$2" "$3" <<'PY'
import json,sys
t,c,i=sys.argv[1:4]
try:
    d=json.loads(open(f'gpurun_out/bench_{t}_{c}_{i}.json').read().strip().splitlines()[-1])
    print(f'BENCH threads={t} cluster={c} inflight={i}: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'fps_sa1 ms', d['kernels'][0]['ms'])
except Exception as e: print('bench parse error', t,c,i, e)
PY
done
MAXCONN=32 timeout 300 python scripts/gpu_timeline.py 4 2>&1 | head -12
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 1 > gpurun_out/ncu_launch.log 2>&1
