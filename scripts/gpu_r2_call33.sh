#!/bin/bash
# round 2, GPU session 33: which levels should use the fused Swin attention kernel? (RS_SWIN_FUSE_MIN_PAIRS), batch 16 / 8 / 1
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s33
for B in 16 1 8; do
  for T in 1 9 33 129 100000; do
    RS_SWIN_FUSE_MIN_PAIRS=$T timeout 300 python bench.py --quick --steps 8 --batch $B 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $B min_pairs $T:', round(d['ms_per_denoise_step'],4), 'ms/step launches', d['launches_per_denoise_step'])" >> $O/${S}_min_pairs.log
  done
done
cat $O/${S}_min_pairs.log
