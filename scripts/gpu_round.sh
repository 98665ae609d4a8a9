#!/bin/bash
# GPU session 41: MLP chunk-order rotation (spread the weight-tile requests of concurrent pairs)
mkdir -p gpurun_out
for rot in 0 1; do echo "rotate=$rot"; RS_MLP_ROTATE=$rot timeout 300 python scripts/mlp_timeline.py 2>&1 | head -1; RS_MLP_ROTATE=$rot timeout 300 python scripts/mlp_timeline.py 16 32 32 2>&1 | head -1; done
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "mlp" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_unet.py -q 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b16.log 2> gpurun_out/bench_b16.err; cat gpurun_out/bench_b16.log | cut -c1-420
