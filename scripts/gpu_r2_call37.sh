#!/bin/bash
# round 2, GPU session 37: compute-sanitizer (memcheck, racecheck) over the tcgen05 Swin attention unit cases
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s37
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout=800 -k "swin_attention_half_fused and tc and not case5" > $O/${S}_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/${S}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout=800 -k "swin_attention_half_fused and tc and case0 and 192" > $O/${S}_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/${S}_racecheck.log
grep -c "Invalid\|Race\|ERROR SUMMARY\|Hazard" $O/${S}_memcheck.log $O/${S}_racecheck.log
tail -8 $O/${S}_memcheck.log | cut -c1-300; tail -25 $O/${S}_racecheck.log | cut -c1-300
