#!/bin/bash
# GPU session 42 (2 GPUs): the driver's N=2 launch with the round-1 end-state build
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "rc=$?"; cat gpurun_out/bench_n2.log | cut -c1-330; tail -3 gpurun_out/bench_n2.err
