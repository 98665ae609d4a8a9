#!/bin/bash
# round 2, GPU session 3: parity suite with in-drain statistics + concurrent low-resolution batch slices; A/B of the stream count
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s3
timeout 1800 python -m pytest tests -m gpu -q -rA --timeout=600 > $O/${S}_pytest_full.log 2>&1
tail -60 $O/${S}_pytest_full.log > $O/${S}_pytest.log
timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_default.log 2>$O/${S}_quick_default.err
RS_LOWRES_STREAMS=1 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_streams1.log 2>/dev/null
RS_LOWRES_STREAMS=4 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_streams4.log 2>/dev/null
RS_LOWRES_STREAMS=4 RS_LOWRES_TILES=128 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_streams4_tiles128.log 2>/dev/null
RS_LOWRES_STREAMS=2 RS_LOWRES_TILES=128 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_streams2_tiles128.log 2>/dev/null
RS_LOWRES_STREAMS=2 RS_LOWRES_TILES=32 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_streams2_tiles32.log 2>/dev/null
RS_LOWRES_STREAMS=1 RS_MLP_NORM_FUSE=1 timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_streams1_mlpnorm.log 2>/dev/null
for k in 1 2 4 32; do
  RS_LOWRES_STREAMS=1 RS_SKIP_KINDS=$k timeout 300 python bench.py --quick --steps 5 > $O/${S}_quick_streams1_skip$k.log 2>/dev/null
done
RS_LOWRES_STREAMS=1 timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
echo done > $O/${S}_done.txt
