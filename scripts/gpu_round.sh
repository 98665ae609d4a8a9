#!/bin/bash
# GPU session 4: staged epilogue + fused GN statistics: op tests, model parity, timelines, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider -x > gpurun_out/ops.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -q -s -p no:cacheprovider > gpurun_out/unet.log 2>&1
timeout 300 python scripts/conv_timeline.py > gpurun_out/tl_default.log 2>&1
timeout 400 python bench.py --steps 3 > gpurun_out/bench.log 2>&1
RS_GN_FUSE=0 timeout 400 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_nofuse.log 2>&1
tail -15 gpurun_out/ops.log
grep -E "parity|property|passed|failed|Error" gpurun_out/unet.log | tail -30
grep -E "^---|us/launch" gpurun_out/tl_default.log
tail -c 1800 gpurun_out/bench.log
tail -c 600 gpurun_out/bench_nofuse.log
