#!/bin/bash
# GPU session 8: templated cg1/cg2 kernels, attention v2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider > gpurun_out/tests_ops.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -q -p no:cacheprovider > gpurun_out/tests_unet.log 2>&1
timeout 300 python scripts/conv_timeline.py > gpurun_out/tl_default.log 2>&1
timeout 300 python scripts/profile_ops.py 16 > gpurun_out/ops_b16.log 2>&1
timeout 400 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -6 gpurun_out/tests_ops.log
tail -6 gpurun_out/tests_unet.log
grep -E "^---|us/launch" gpurun_out/tl_default.log
head -30 gpurun_out/ops_b16.log
tail -c 600 gpurun_out/bench.log
