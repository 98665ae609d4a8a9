#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_other_tasks.py 2>&1 | tee gpurun_out/bench_other_tasks.log | tail -4
