#!/bin/bash
# round 2, GPU session 31 (2 GPUs), fused tcgen05 Swin attention on by default: two-GPU sampler test, bench --gpus 2 (shard parity, NCCL uint8 gather of the bookend path)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s31
nvidia-smi -L > $O/${S}_gpus.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA --timeout=800 > $O/${S}_pytest_multi.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 --no-library-baseline --no-other-configs > $O/${S}_bench_n2.log 2> $O/${S}_bench_n2.err
echo done > $O/${S}_done.txt
tail -4 $O/${S}_pytest_multi.log; head -c 900 $O/${S}_bench_n2.log; echo
