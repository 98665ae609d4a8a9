#!/bin/bash
# One GPU session: diagnostics, operator tests (per implementation, isolated processes), model parity, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python scripts/gpu_diag.py > gpurun_out/diag.log 2>&1
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "not tcgen05" -p no:cacheprovider > gpurun_out/ops_other.log 2>&1
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "tcgen05" -p no:cacheprovider > gpurun_out/ops_tcgen05.log 2>&1
RS_CONV_IMPL=simt RS_ATTN_IMPL=simt timeout 1200 python -m pytest tests/test_gpu_unet.py -q -s -p no:cacheprovider -k "tiny or fresh or rect" > gpurun_out/unet_simt.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_unet.py -q -s -p no:cacheprovider > gpurun_out/unet.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/diag.log gpurun_out/ops_other.log gpurun_out/ops_tcgen05.log gpurun_out/unet_simt.log gpurun_out/unet.log gpurun_out/bench.log
