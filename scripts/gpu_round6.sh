#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_timeline.py 1 > gpurun_out/timeline_1.log 2>&1; cat gpurun_out/timeline_1.log
timeout 300 python scripts/gpu_timeline.py 4 > gpurun_out/timeline_4.log 2>&1; cat gpurun_out/timeline_4.log
timeout 300 python scripts/gpu_timeline.py 8 512 4 > gpurun_out/timeline_8.log 2>&1; cat gpurun_out/timeline_8.log
timeout 300 python -m pytest tests/test_gpu_engine.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -3
