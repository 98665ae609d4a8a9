#!/bin/bash
# timing of the fused Swin attention op alone + the quick in-graph figure (A/B runs of kernel variants)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=${1:-r2_sx}
timeout 300 python scripts/swin_tc_diag.py time > $O/${S}_swin_tc_time.log 2>&1
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>/dev/null
grep "time impl=tc" $O/${S}_swin_tc_time.log; grep "tile 1 workers" $O/${S}_swin_tc_time.log | cut -c1-260; head -c 200 $O/${S}_quick_default.log; echo
