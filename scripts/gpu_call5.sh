#!/bin/bash
# session call 5: suspend-time hint of the mbarrier waits / polling MMA issuers (A/B through tuning knobs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dense or forward or concurrent" 2>&1 | tail -n 4
for tune in "" "sa_wait_ns=100" "sa_wait_ns=400" "sa_variant=3" "sa_variant=3,sa_wait_ns=100"; do
  echo "--- stress VNB_TUNE='$tune'"; VNB_TUNE="$tune" timeout 200 python scripts/gpu_stress.py 12 200 sa1,sa2,sa3,sa4,prop,fp_vote_fused 2>&1 | tail -n 6
done
echo "--- sa1 trace (sa_wait_ns=100)"; VNB_TUNE="sa_wait_ns=100" timeout 200 python scripts/gpu_trace_sa1.py 2>&1 | tail -n 10
echo "--- sa2 trace (sa_wait_ns=100)"; VNB_TUNE="sa_wait_ns=100" timeout 200 python scripts/gpu_trace_sa2.py 1 2>&1 | tail -n 8
echo "--- sa2 trace (sa_variant=3,sa_wait_ns=100)"; VNB_TUNE="sa_variant=3,sa_wait_ns=100" timeout 200 python scripts/gpu_trace_sa2.py 1 2>&1 | tail -n 8
for tune in "sa_wait_ns=100" "sa_variant=3,sa_wait_ns=100"; do
  t1=$(echo $tune | sed 's/,/ --tune /g')
  python bench.py --no-cpu-baseline --tune $t1 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "bench [$tune] exit=$?"
  python - <<'PY'
import json
d = json.loads(open('gpurun_out/c5_bench.json').read().strip().splitlines()[-1])
print('   value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'lat1', d.get('latency_ms_inflight1'))
PY
done
