#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; python bench.py --steps 600 --warmup 32 --no-cpu-baseline "$@" 2>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'enq', round(d['host_enqueue_ms_per_step']['device_loop'],3))"; grep -v Warning gpurun_out/err.txt | tail -n 2; }
run --inflight 8
run --inflight 12
run --inflight 16
run --inflight 24
run --inflight 32
CUDA_DEVICE_MAX_CONNECTIONS=8 run --inflight 16
