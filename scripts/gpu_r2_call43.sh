#!/bin/bash
# round 2, GPU session 43: few-slot GroupNorm combines request every pair at once — per-op table + quick bench + GN tests
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s43
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -m gpu -q --timeout=600 -x -k "groupnorm or statistics or norm2 or golden or batch_independence" > $O/${S}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${S}_pytest.log
timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>/dev/null
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default_b.log 2>/dev/null
tail -3 $O/${S}_pytest.log; grep " gn " $O/${S}_per_op_table_b16.log | head -12; head -c 160 $O/${S}_quick_default.log; echo; head -c 160 $O/${S}_quick_default_b.log; echo
