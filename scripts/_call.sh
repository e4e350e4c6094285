timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/gpu_peer_test.py 2>&1 | grep -v "Warning\|warn\|^\*\|OMP" | tail -n 5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu_driverlike.json 2> gpurun_out/r2_bench_2gpu_driverlike.err; echo "bench exit=$?"; grep -v "Warn\|warn\|^\*\|OMP" gpurun_out/r2_bench_2gpu_driverlike.err | tail -3
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu_driverlike.json').read().strip().splitlines()[-1])
print('2gpu driver-like: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), d['config']['parallelism'][-20:], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 4 --warmup 3 2>/dev/null | tail -1 | cut -c1-300
