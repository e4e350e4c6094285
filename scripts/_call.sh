timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_peer_diag.py 2>&1 | grep "rank \|Error:" | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/gpu_peer_test.py 2>&1 | grep -v "Warning\|warn\|^\*\|OMP" | tail -n 8
for T in peer; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --transport $T > gpurun_out/r2_bench_2gpu_$T.json 2> gpurun_out/r2_bench_2gpu_$T.err; echo "bench $T exit=$?"; grep -v "Warn\|warn\|^\*\|OMP" gpurun_out/r2_bench_2gpu_$T.err | tail -3
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu_$T.json').read().strip().splitlines()[-1])
print('$T 2gpu: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), d['config']['parallelism'])
PY
done
