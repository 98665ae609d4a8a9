#!/bin/bash
# round 2, GPU session 39: memcheck over the multi-tile unit case and the plan-level fused-vs-four-launch test
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s39
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=580 -k "swin_attention_half_fused and tc and case5" > $O/${S}_memcheck_case5.log 2>&1
echo "rc=$?" >> $O/${S}_memcheck_case5.log
RS_SWIN_FUSE_MIN_PAIRS=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout=880 -k "fused_swin_attention_plan" > $O/${S}_memcheck_plan.log 2>&1
echo "rc=$?" >> $O/${S}_memcheck_plan.log
tail -4 $O/${S}_memcheck_case5.log | cut -c1-200; tail -4 $O/${S}_memcheck_plan.log | cut -c1-200
