#!/bin/bash
# round 2, GPU session 20: seventh version (relative-position bias as a 225-entry table per head in shared memory)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s20
timeout 300 python scripts/swin_tc_diag.py time > $O/${S}_swin_tc_time.log 2>&1
echo "time rc=$?" >> $O/${S}_swin_tc_time.log
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "swin_attention_half_fused" > $O/${S}_pytest_swin.log 2>&1
RS_SWIN_FUSE=1 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_swinfuse_tc.log 2>$O/${S}_quick_swinfuse_tc.err
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>/dev/null
tail -16 $O/${S}_swin_tc_time.log; tail -3 $O/${S}_pytest_swin.log; for f in swinfuse_tc default; do head -c 330 $O/${S}_quick_$f.log; echo; done
echo done > $O/${S}_done.txt
