#!/bin/bash
# round 2, GPU session 7: default back to the four-launch attention half; VQ-GAN bookend per-op tables; full bench incl. bookends
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s7
timeout 600 python scripts/profile_vq.py > $O/${S}_vq_per_op_b16.log 2>&1
timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/${S}_bench_b16.log 2> $O/${S}_bench_b16.err
timeout 900 python -m pytest tests/test_gpu_vq.py tests/test_gpu_unet.py -m gpu -q -x --timeout=600 > $O/${S}_pytest.log 2>&1
echo done > $O/${S}_done.txt
