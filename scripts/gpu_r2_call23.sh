#!/bin/bash
# round 2, GPU session 23: why does the batch-1 plan with the fused Swin attention differ? block probes, fused vs unfused
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s23
timeout 600 python scripts/swin_plan_diag.py 1 2 > $O/${S}_plan_diag.log 2>&1
grep -n "N=\|<<<<" $O/${S}_plan_diag.log | head -60
