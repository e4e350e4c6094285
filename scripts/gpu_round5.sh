#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_index_ops test_gpu_dense test_gpu_engine; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -x --no-header -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 12 gpurun_out/$f.log | cut -c1-300
done
timeout 600 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe exit=$?" >> gpurun_out/summary.txt
head -n 24 gpurun_out/probe.log
for cfg in "256 8 4" "256 8 8" "512 4 4" "512 4 8" "1024 2 8" "1024 2 12" "512 2 12"; do
set -- $cfg
timeout 600 python bench.py --steps 200 --warmup 10 --fps-threads $1 --fps-cluster $2 --inflight $3 --no-cpu-baseline > gpurun_out/bench_$1_$2_$3.json 2> gpurun_out/bench_$1_$2_$3.err; echo "bench $cfg exit=$?" >> gpurun_out/summary.txt
python - "$1" "$2" "$3" <<'PY'
import json,sys
t,c,i=sys.argv[1:4]
try:
    d=json.loads(open(f'gpurun_out/bench_{t}_{c}_{i}.json').read().strip().splitlines()[-1])
    print(f'BENCH threads={t} cluster={c} inflight={i}: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'fps_sa1 ms', d['kernels'][0]['ms'])
except Exception as e: print('bench parse error', t,c,i, e)
PY
done
cat gpurun_out/summary.txt
