#!/bin/bash
# round 2, GPU session 41: memcheck over whole-model tests (full-width forward golden, 15-step loop golden, batch independence)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s41
timeout 540 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout=520 -k "(forward_vs_reference_golden and realsr) or loop_realsr_15 or batch_independence" > $O/${S}_memcheck_unet.log 2>&1
echo "rc=$?" >> $O/${S}_memcheck_unet.log
grep -c "Invalid\|Error:" $O/${S}_memcheck_unet.log; tail -8 $O/${S}_memcheck_unet.log | cut -c1-220
