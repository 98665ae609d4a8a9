#!/bin/bash
# round 2, GPU session 44 (2 GPUs): bench --gpus 2 with the end-state defaults
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s44
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29644 bench.py --gpus 2 --steps 5 --warmup 3 --no-library-baseline --no-other-configs --no-cpu-baseline > $O/${S}_bench_n2.log 2> $O/${S}_bench_n2.err
head -c 1000 $O/${S}_bench_n2.log; echo
