#!/bin/bash
# round 2, GPU session 14: second version (ILP, cp.async gather, cheaper statistics, polling issuer) of the tcgen05 Swin attention kernel (swin_attn_tc.cuh)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s14
timeout 300 python scripts/swin_tc_diag.py diag > $O/${S}_swin_tc_diag.log 2>&1
echo "diag rc=$?" >> $O/${S}_swin_tc_diag.log
timeout 300 python scripts/swin_tc_diag.py time > $O/${S}_swin_tc_time.log 2>&1
echo "time rc=$?" >> $O/${S}_swin_tc_time.log
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "swin_attention_half_fused" > $O/${S}_pytest_swin.log 2>&1
RS_SWIN_FUSE=1 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_swinfuse_tc.log 2>$O/${S}_quick_swinfuse_tc.err
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>/dev/null
tail -5 $O/${S}_swin_tc_diag.log; tail -12 $O/${S}_swin_tc_time.log; tail -5 $O/${S}_pytest_swin.log; cat $O/${S}_quick_swinfuse_tc.log | head -c 600; echo; cat $O/${S}_quick_default.log | head -c 600
echo done > $O/${S}_done.txt
