#!/bin/bash
# round 2, GPU session 21: ncu --set full of the tcgen05 Swin attention kernel, seventh version
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s21
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swin_attn_tc -c 1 -o $O/${S}_swin_tc --force-overwrite python scripts/swin_tc_diag.py ncu > $O/${S}_ncu.log 2>&1
ls -la $O/${S}_swin_tc.ncu-rep
tail -3 $O/${S}_ncu.log
