#!/bin/bash
# round 2, GPU session 2: full parity suite (incl. VQ-GAN bookends, edges, tiling) + A/B / ablation of the step
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s2
timeout 1800 python -m pytest tests -m gpu -q -rA --timeout=600 2>&1 | tail -260 > $O/${S}_pytest.log
timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_default.log 2>$O/${S}_quick_default.err
RS_MLP_NORM_FUSE=0 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_nomlpnorm.log 2>/dev/null
for k in 1 2 4 32; do
  RS_SKIP_KINDS=$k timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_skip$k.log 2>/dev/null
done
timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
echo done > $O/${S}_done.txt
