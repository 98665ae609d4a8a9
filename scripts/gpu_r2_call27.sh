#!/bin/bash
# round 2, GPU session 27: next group's Q / K drain under the PV wait
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s27
timeout 300 python scripts/swin_tc_diag.py time > $O/${S}_swin_tc_time.log 2>&1
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -m gpu -q --timeout=600 -x -k "swin or batch_independence or golden or loop" > $O/${S}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${S}_pytest.log
timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_default.log 2>$O/${S}_quick_default.err
RS_SWIN_FUSE=0 timeout 300 python bench.py --quick --steps 8 > $O/${S}_quick_noswinfuse.log 2>/dev/null
tail -12 $O/${S}_swin_tc_time.log | cut -c1-250; tail -4 $O/${S}_pytest.log; for f in default noswinfuse; do head -c 300 $O/${S}_quick_$f.log; echo; done
echo done > $O/${S}_done.txt
