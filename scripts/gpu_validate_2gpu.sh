#!/bin/bash
# Multi-GPU part of the validation (gpurun --gpus 2 -- bash scripts/gpu_validate_2gpu.sh): the peer-memory exchange against
# NCCL and a host merge, then the weak-scaling bench lines with both transports.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29513 scripts/gpu_peer_test.py 2>&1 | grep -v "Warning\|warn\|^\*\|OMP" | tail -n 5
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu_driverlike.json 2> gpurun_out/r2_bench_2gpu_driverlike.err; echo "bench (driver-like) exit=$?"
timeout 300 $TR --master-port 29515 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --transport peer > gpurun_out/r2_bench_2gpu_peer.json 2> gpurun_out/r2_bench_2gpu_peer.err; echo "bench peer exit=$?"
timeout 300 $TR --master-port 29516 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --transport nccl > gpurun_out/r2_bench_2gpu_nccl.json 2> gpurun_out/r2_bench_2gpu_nccl.err; echo "bench nccl exit=$?"
python - <<'PY'
import json
for f in ('r2_bench_2gpu_driverlike', 'r2_bench_2gpu_peer', 'r2_bench_2gpu_nccl'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), d['config']['parallelism'][-40:])
    except Exception as e:
        print(f, 'unreadable', e)
PY
