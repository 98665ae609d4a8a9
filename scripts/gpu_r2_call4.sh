#!/bin/bash
# round 2, GPU session 4: parity suite, full bench line (bookends, other configs, library + CPU baselines), smoke, ncu captures
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s4
timeout 1800 python -m pytest tests -m gpu -q -rA --timeout=600 > $O/${S}_pytest_full.log 2>&1
tail -40 $O/${S}_pytest_full.log > $O/${S}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${S}_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${S}_bench_b16.log 2> $O/${S}_bench_b16.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${S}_bench_reference_arm.log 2>/dev/null
timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file $O/${S}_ncu_launch_list_forward_b16.csv python scripts/profile_forward.py --iters 1 > $O/${S}_ncu_run.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_gemm|mlp_fused" -c 24 \
  -o $O/${S}_gemm_full python scripts/profile_forward.py --iters 1 > $O/${S}_ncu_full_run.log 2>&1
ncu -i $O/${S}_gemm_full.ncu-rep --page raw --csv > $O/${S}_gemm_full_raw.csv 2>/dev/null
echo done > $O/${S}_done.txt
