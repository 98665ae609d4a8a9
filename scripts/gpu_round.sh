#!/bin/bash
# GPU session 13: calibrated split-K; full suite + bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/tests_all.log 2>&1
timeout 300 python scripts/profile_ops.py 16 > gpurun_out/ops_b16.log 2>&1
timeout 500 python bench.py > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --batch 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_b1.log 2>&1
tail -5 gpurun_out/tests_all.log
head -20 gpurun_out/ops_b16.log
grep -o '"ms_per_denoise_step": [0-9.]*' gpurun_out/bench.log gpurun_out/bench_b1.log
grep -o '"value": [0-9.]*' gpurun_out/bench.log | head -3
