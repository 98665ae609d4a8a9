#!/bin/bash
# round 2, GPU session 12: gn_apply with the first rows in flight before the statistics prologue
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s12
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout=600 -k "groupnorm or statistics or gn" > $O/${S}_pytest.log 2>&1
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>$O/${S}_quick_default.err
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default_b.log 2>/dev/null
RS_SKIP_KINDS=4 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_skip4.log 2>/dev/null
timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
timeout 600 python scripts/profile_vq.py > $O/${S}_vq_per_op_b16.log 2>&1
echo done > $O/${S}_done.txt
