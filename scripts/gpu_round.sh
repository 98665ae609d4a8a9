#!/bin/bash
# GPU session 37: fused MLP with the hidden dimension split over two pairs of a cluster for few-tile layers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "mlp or split_k" 2>&1 | tail -4
for hsp in 1 2; do echo "hsplit=$hsp"; RS_MLP_HSPLIT=$hsp timeout 300 python scripts/mlp_timeline.py 16 8 8 2>&1 | head -1; RS_MLP_HSPLIT=$hsp timeout 300 python scripts/mlp_timeline.py 16 16 16 2>&1 | head -1; RS_MLP_HSPLIT=$hsp timeout 300 python scripts/mlp_timeline.py 16 32 32 2>&1 | head -1; done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b16.log 2> gpurun_out/bench_b16.err; cat gpurun_out/bench_b16.log | cut -c1-420
