#!/bin/bash
# round 2, GPU session 28: the tcgen05 Swin kernel finalises its output statistics (run-length arrival per image)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s28
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -x > $O/${S}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${S}_pytest.log
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>$O/${S}_quick_default.err
RS_SWIN_FINALIZE=0 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_nofinalize.log 2>/dev/null
RS_MLP_NORM_FUSE=1 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_mlpnorm.log 2>/dev/null
RS_SWIN_FUSE=0 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_noswinfuse.log 2>/dev/null
RS_MLP_NORM_FUSE=1 timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_mlpnorm_b16.log 2>&1
tail -6 $O/${S}_pytest.log; for f in default nofinalize mlpnorm noswinfuse; do head -c 330 $O/${S}_quick_$f.log; echo; done
echo done > $O/${S}_done.txt
