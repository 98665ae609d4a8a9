#!/bin/bash
# GPU session 36: cluster split-K restricted to S = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -x -q 2>&1 | tail -4
timeout 300 python scripts/profile_ops.py > gpurun_out/per_op.log 2>&1; head -14 gpurun_out/per_op.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b16.log 2> gpurun_out/bench_b16.err; cat gpurun_out/bench_b16.log | cut -c1-420
