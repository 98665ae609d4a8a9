#!/bin/bash
# round 2, GPU session 5: fused Swin attention half (unit + model parity), A/B against the four-launch sequence, full bench
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s5
timeout 1800 python -m pytest tests -m gpu -q -rA --timeout=600 > $O/${S}_pytest_full.log 2>&1
tail -40 $O/${S}_pytest_full.log > $O/${S}_pytest.log
timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_default.log 2>$O/${S}_quick_default.err
RS_SWIN_FUSE=0 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_noswinfuse.log 2>/dev/null
RS_SKIP_KINDS=8 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_skip8.log 2>/dev/null
RS_SKIP_KINDS=4 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_skip4.log 2>/dev/null
timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${S}_bench_b16.log 2> $O/${S}_bench_b16.err
echo done > $O/${S}_done.txt
