#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s6
timeout 300 python scripts/profile_swin.py > $O/${S}_swin_times.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swin_attn -c 6 -o $O/${S}_swin_full python scripts/profile_swin.py > $O/${S}_ncu_run.log 2>&1
ncu -i $O/${S}_swin_full.ncu-rep --page raw --csv > $O/${S}_swin_raw.csv 2>/dev/null
ncu -i $O/${S}_swin_full.ncu-rep --page source --csv --kernel-id :::2 > $O/${S}_swin_source_64.csv 2>/dev/null
echo done > $O/${S}_done.txt
