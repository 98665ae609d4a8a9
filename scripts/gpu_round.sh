#!/bin/bash
# GPU session 40: validation of the round-1 end state + refreshed ncu artifacts
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.log | cut -c1-250
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 1 > gpurun_out/bench_b1.log 2> gpurun_out/bench_b1.err; cat gpurun_out/bench_b1.log | cut -c1-250
timeout 300 python scripts/profile_ops.py > gpurun_out/per_op.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_forward.csv python scripts/profile_forward.py --iters 1 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_gemm|mlp_fused" -c 20 \
    -o gpurun_out/prof_gemm2 -f python scripts/profile_forward.py --iters 1 > gpurun_out/ncu_full.log 2>&1
du -sm gpurun_out; ls gpurun_out
