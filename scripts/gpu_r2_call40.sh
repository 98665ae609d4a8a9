#!/bin/bash
# round 2, GPU session 40: memcheck over the whole operator test file (every kernel of the library through the C ABI)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s40
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=880 > $O/${S}_memcheck_ops.log 2>&1
echo "rc=$?" >> $O/${S}_memcheck_ops.log
grep -c "Invalid\|Error:" $O/${S}_memcheck_ops.log; tail -8 $O/${S}_memcheck_ops.log | cut -c1-220
