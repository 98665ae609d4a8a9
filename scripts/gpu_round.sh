#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/mlp_timeline.py > gpurun_out/mlp_tl.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider -k "mlp or epilogue" > gpurun_out/tests_new.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -q -p no:cacheprovider > gpurun_out/tests_unet.log 2>&1
timeout 300 python scripts/profile_ops.py 16 > gpurun_out/ops_b16.log 2>&1
timeout 400 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
cat gpurun_out/mlp_tl.log
tail -4 gpurun_out/tests_new.log
tail -4 gpurun_out/tests_unet.log
head -24 gpurun_out/ops_b16.log
tail -c 500 gpurun_out/bench.log
