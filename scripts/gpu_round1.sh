#!/bin/bash
# First GPU bring-up: each test file in its own process (a trapped kernel poisons the CUDA context), bounded by timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in test_gpu_index_ops test_gpu_ref_kernels test_gpu_dense test_gpu_forward; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu --timeout 300 -x --no-header -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 30 gpurun_out/$f.log
done
timeout 300 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/probe.log
cat gpurun_out/summary.txt
