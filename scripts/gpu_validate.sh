#!/bin/bash
# One-stop GPU validation (run on a B200 box, e.g.  gpurun --timeout 2400 -- bash scripts/gpu_validate.sh):
# parity tests, smoke, the default bench line (both arms), the per-stage cost table under load, and the ncu evidence
# that profiles/ keeps (launch list of the bench command + one --set full capture of the top kernels).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -n 2
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench exit=$?"
timeout 600 python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/bench_reference.json 2>/dev/null
timeout 300 python scripts/gpu_stress.py 12 200 > gpurun_out/stress.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 2 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"fps_pruned_kernel|sa_ws2_kernel|sa1_ws2_kernel|fp_chain_kernel|grid_query_kernel|nms_clip_kernel|linear_tc_kernel" -c 16 -f \
    -o gpurun_out/full python scripts/gpu_one_forward.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1),
      'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 2), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
for k in d['kernels']:
    print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
