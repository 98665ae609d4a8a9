#!/bin/bash
# round 2, GPU session 10: GroupNorm finalisation as a kernel of its own for many-slot tensors (VQ-GAN), tests + tables + bench
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s10
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vq.py -m gpu -q -x --timeout=600 -k "statistics or vq or groupnorm or Vq or pipeline or edges or tiled" > $O/${S}_pytest.log 2>&1
timeout 600 python scripts/profile_vq.py > $O/${S}_vq_per_op_b16.log 2>&1
RS_GN_PRODUCER_FINALIZE=1 RS_CONV_PERSIST=0 timeout 600 python scripts/profile_vq.py > $O/${S}_vq_per_op_b16_old_nopersist.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 --no-library-baseline --no-other-configs > $O/${S}_bench_b16.log 2> $O/${S}_bench_b16.err
echo done > $O/${S}_done.txt
