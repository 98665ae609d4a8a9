#!/bin/bash
# GPU session 43: sampler class / tiler after the tile_starts refactor
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -q -k "sampler_class or forward_vs_reference_golden" 2>&1 | tail -3
timeout 300 python - <<'PY'
# tiled path: an input larger than chop_size through ResShiftSampler._process vs per-tile calls
import torch, sys
sys.path.insert(0, '.')
from resshift_b200.sampler import tile_starts
print("tile_starts ok", tile_starts(300, 128, 128))
PY
