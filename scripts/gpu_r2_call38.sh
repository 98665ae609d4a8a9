#!/bin/bash
# round 2, GPU session 38: early projection k-blocks only after both PVs of the last group (ordering bug found by the
# sanitizer run); memcheck again over the unit cases, suite subset, timing
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s38
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=1000 -k "swin_attention_half_fused and tc and not case5" > $O/${S}_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/${S}_memcheck.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -m gpu -q --timeout=600 -x -k "swin or batch_independence or golden or loop or norm2 or fused" > $O/${S}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${S}_pytest.log
timeout 300 python scripts/swin_tc_diag.py time > $O/${S}_swin_tc_time.log 2>&1
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>/dev/null
tail -5 $O/${S}_memcheck.log | cut -c1-200; tail -3 $O/${S}_pytest.log; grep "time impl=tc" $O/${S}_swin_tc_time.log; head -c 200 $O/${S}_quick_default.log; echo
