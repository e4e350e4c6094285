#!/bin/bash
# bench + ncu launch list of the bench command + one --set full capture of the top kernels
mkdir -p gpurun_out
python bench.py --steps 400 --warmup 16 > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; echo "bench exit=$?"; tail -n 3 gpurun_out/bench_s3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline'], 'clocks', d['clocks'], 'enq', d['host_enqueue_ms_per_step'])
for k in d['kernels']: print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fps_pruned_kernel|sa_ws2_kernel|sa1_ws2_kernel|linear_tc_kernel|nms_mask_kernel|grid_query_kernel" -c 24 -f -o gpurun_out/full_s3 python scripts/gpu_one_forward.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit=$?"; tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
