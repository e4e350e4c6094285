#!/bin/bash
# One-stop GPU validation (run on a B200 box, e.g.  gpurun --timeout 2400 -- bash scripts/gpu_validate.sh):
# parity tests, smoke, the bench lines (default, driver-like, reference arm), the per-stage cost table under load, the
# sanitizers, and the ncu evidence that profiles/ keeps (launch list of the bench command + one --set full capture).
mkdir -p gpurun_out
rm -f gpurun_out/stage_errors.json
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -n 2
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench exit=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu_driverlike.json 2> gpurun_out/r2_bench_1gpu_driverlike.err; echo "bench (driver-like) exit=$?"
timeout 600 python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "reference arm exit=$?"
timeout 300 python scripts/gpu_stress.py 12 200 three_nn,prop_rest,bq34,sa1,sa2,sa3,sa4,prop,bq1,bq2,fps_nested,fp_vote_fused,nms,fps1 > gpurun_out/r2_stress.txt 2>&1
timeout 200 python scripts/gpu_fps_phases.py room > gpurun_out/r2_fps_phases.txt 2>&1
timeout 200 python scripts/gpu_fps_phases.py uniform | head -n 2 >> gpurun_out/r2_fps_phases.txt 2>&1
timeout 200 python scripts/gpu_nms_probe.py > gpurun_out/r2_nms_probe.txt 2>&1
# stage timelines of CTA 0 of the three fused tensor-core kernels (debug trace)
( timeout 200 python scripts/gpu_trace_sa1.py; timeout 200 python scripts/gpu_trace_sa2.py 1; timeout 200 python scripts/gpu_trace_fp.py ) > gpurun_out/r2_stage_traces.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_ncu_launch_list_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 2 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"fps_pruned_kernel|sa_ws2_kernel|sa1_ws2_kernel|fp_chain_kernel|grid_query2_kernel|group_a0_kernel|group_rel_kernel|nms_cloud_kernel|linear_tc_kernel" -c 22 -f \
    -o gpurun_out/r2_full python scripts/gpu_one_forward.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit=$?"
python - <<'PY'
import json
for f in ('r2_bench_1gpu', 'r2_bench_1gpu_driverlike'):
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1),
          'cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['cores']), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
d = json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
for k in d['kernels']:
    print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
bash scripts/gpu_sanitize.sh r2 > gpurun_out/r2_sanitizer_summary.txt 2>&1; tail -n 30 gpurun_out/r2_sanitizer_summary.txt
