#!/bin/bash
# round 2, GPU session 35: window-granularity split over the CTAs, idle half of a one-window tile skips its work
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s35
timeout 300 python scripts/swin_tc_diag.py time > $O/${S}_swin_tc_time.log 2>&1
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -m gpu -q --timeout=600 -x -k "swin or batch_independence or golden or loop or norm2 or fused" > $O/${S}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${S}_pytest.log
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>/dev/null
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default_b.log 2>/dev/null
grep "time impl=tc" $O/${S}_swin_tc_time.log; grep "tile 1 workers" $O/${S}_swin_tc_time.log | cut -c1-260; tail -3 $O/${S}_pytest.log; head -c 200 $O/${S}_quick_default.log; echo; head -c 200 $O/${S}_quick_default_b.log; echo
