#!/bin/bash
# GPU session 25: warp-uniform control loops (elect-predicated tcgen05.mma / commit / TMA issue) in conv + MLP kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -5
timeout 300 python scripts/mlp_timeline.py > gpurun_out/mlp_timeline.log 2>&1; tail -10 gpurun_out/mlp_timeline.log
timeout 300 python scripts/mlp_timeline.py 16 8 8 2>&1 | head -1
timeout 300 python scripts/mlp_timeline.py 16 32 32 2>&1 | head -1
timeout 300 python scripts/conv_timeline.py > gpurun_out/conv_timeline.log 2>&1; cat gpurun_out/conv_timeline.log
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q 2>&1 | tail -5
timeout 300 python scripts/profile_ops.py > gpurun_out/per_op.log 2>&1; head -40 gpurun_out/per_op.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b16.log 2> gpurun_out/bench_b16.err; cat gpurun_out/bench_b16.log | cut -c1-420
