#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_index_ops test_gpu_dense test_gpu_engine test_gpu_forward test_gpu_ref_kernels; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -x --no-header -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 12 gpurun_out/$f.log | cut -c1-300
done
timeout 600 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/probe.log
timeout 300 python scripts/gpu_fps_prof.py 2>&1 | tee gpurun_out/fps_prof.log
for cfg in "256 8 8" "256 4 8" "512 4 8"; do
set -- $cfg
timeout 600 python bench.py --steps 200 --warmup 10 --fps-threads $1 --fps-cluster $2 --inflight $3 --no-cpu-baseline > gpurun_out/bench_$1_$2_$3.json 2> gpurun_out/bench_$1_$2_$3.err; echo "bench $cfg exit=$?" >> gpurun_out/summary.txt
python - "$1" "$2" "$3" <<'PY'
import json,sys
t,c,i=sys.argv[1:4]
try:
    d=json.loads(open(f'gpurun_out/bench_{t}_{c}_{i}.json').read().strip().splitlines()[-1])
    print(f'BENCH threads={t} cluster={c} inflight={i}: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
    if c=='8':
        for k in d['kernels']: print('  ', k['kernel'], k['ms'], round(k['frac'],3))
except Exception as e: print('bench parse error', t,c,i, e)
PY
done
timeout 300 python scripts/gpu_timeline.py 1 2>&1 | head -8
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 1 > gpurun_out/ncu_launch.log 2>&1
for k in sa_ws_kernel sa1_ws_kernel; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 1 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
