#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -q -k "forward_vs_reference_golden" 2>&1 | tail -4
