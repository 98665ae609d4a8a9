#!/bin/bash
# GPU session 29: persistent conv chosen by the cost model (BN re-picked for the epilogue-bound 1x1 layers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q 2>&1 | tail -4
timeout 300 python scripts/profile_ops.py > gpurun_out/per_op.log 2>&1; head -30 gpurun_out/per_op.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b16.log 2> gpurun_out/bench_b16.err; cat gpurun_out/bench_b16.log | cut -c1-420
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 1 > gpurun_out/bench_b1.log 2> gpurun_out/bench_b1.err; cat gpurun_out/bench_b1.log | cut -c1-420
