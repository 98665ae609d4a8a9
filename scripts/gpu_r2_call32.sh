#!/bin/bash
# round 2, GPU session 32 (2 GPUs): two-GPU sampler test with 64x64 latents (the tiny preset has the shipped topology: 4 levels)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s32
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA --timeout=800 > $O/${S}_pytest_multi.log 2>&1
tail -6 $O/${S}_pytest_multi.log | cut -c1-600
