#!/bin/bash
# round 2, GPU session 24: fused vs unfused Swin attention with workspace reuse, batch 1 / 2 / 3, switches
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s24
timeout 600 python scripts/swin_plan_ab.py 1 2 3 > $O/${S}_plan_ab.log 2>&1
cat $O/${S}_plan_ab.log | tail -20
