#!/bin/bash
# session 3, experiment 1: where does the 0.58 ms/step go?  FPS interference vs serial sum of the wide kernels
mkdir -p gpurun_out
run() { echo "== $*"; python bench.py --steps 400 --warmup 16 --no-cpu-baseline "$@" 2>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"; tail -n 3 gpurun_out/err.txt; }
run
run --debug-skip-fps1
run --tune sa_sms=132
run --tune sa_sms=124
run --tune sa_sms=116
run --tune sa_split=2
run --tune sa_split=4
run --tune sa_sms=124 --tune sa_split=2
run --inflight 4
run --inflight 12
python scripts/gpu_timeline.py 8 2>&1 | tail -n 30
