#!/bin/bash
# round 2, GPU session 25: the failing batch-16-then-batch-1 sequence under switches
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s25
timeout 600 python scripts/swin_plan_seq.py > $O/${S}_plan_seq.log 2>&1
cat $O/${S}_plan_seq.log | tail -20
