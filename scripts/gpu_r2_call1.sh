#!/bin/bash
# round 2, GPU session 1: parity suite, bench line, A/B runs, kernel-family ablation inside the graph,
# warm-cache ncu launch list, per-op table
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${S}_smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -rA --timeout=600 2>&1 | tail -150 > $O/${S}_pytest.log
timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_default.log 2>$O/${S}_quick_default.err
RS_CONV_TAILSKIP=0 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_notailskip.log 2>/dev/null
RS_MLP_NORM_FUSE=0 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_nomlpnorm.log 2>/dev/null
for k in 1 2 4 8 16 32 6 46 47; do
  RS_SKIP_KINDS=$k timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_skip$k.log 2>/dev/null
done
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${S}_bench_b16.log 2> $O/${S}_bench_b16.err
timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file $O/${S}_ncu_launch_list_warm_forward_b16.csv python scripts/profile_forward.py --iters 1 > $O/${S}_ncu_run.log 2>&1
echo done > $O/${S}_done.txt
