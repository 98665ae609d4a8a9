#!/bin/bash
# round 2, GPU session 17: tcgen05 Swin attention, fourth version (real LDS/STS, test_wait polling, cheaper statistics / affine)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s17
timeout 300 python scripts/swin_tc_diag.py diag > $O/${S}_swin_tc_diag.log 2>&1
echo "diag rc=$?" >> $O/${S}_swin_tc_diag.log
timeout 300 python scripts/swin_tc_diag.py time > $O/${S}_swin_tc_time.log 2>&1
echo "time rc=$?" >> $O/${S}_swin_tc_time.log
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "swin_attention_half_fused" > $O/${S}_pytest_swin.log 2>&1
RS_SWIN_FUSE=1 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_swinfuse_tc.log 2>$O/${S}_quick_swinfuse_tc.err
RS_SWIN_FUSE=1 RS_MLP_NORM_FUSE=1 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_swinfuse_tc_mlpnorm.log 2>/dev/null
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>/dev/null
RS_SWIN_FUSE=1 timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_swinfuse_b16.log 2>&1
grep "max=" $O/${S}_swin_tc_diag.log | head -12; tail -14 $O/${S}_swin_tc_time.log; tail -3 $O/${S}_pytest_swin.log; for f in swinfuse_tc swinfuse_tc_mlpnorm default; do head -c 330 $O/${S}_quick_$f.log; echo; done
grep -i "swin_attn\|^ops" $O/${S}_per_op_table_swinfuse_b16.log | head
echo done > $O/${S}_done.txt
