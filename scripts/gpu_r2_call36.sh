#!/bin/bash
# round 2, GPU session 36: end-of-round state (fused Swin attention on >= 96-pair levels) — whole GPU suite, smoke, full bench line + reference arm, per-op tables, ncu
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s36
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout=600 > $O/${S}_pytest_full.log 2>&1
tail -40 $O/${S}_pytest_full.log > $O/${S}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${S}_smoke.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 > $O/${S}_bench_b16.log 2> $O/${S}_bench_b16.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${S}_bench_reference_arm.log 2>/dev/null
timeout 300 python scripts/profile_ops.py > $O/${S}_per_op_table_b16.log 2>&1
timeout 300 python scripts/swin_tc_diag.py time > $O/${S}_swin_tc_time.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file $O/${S}_ncu_launch_list_forward_b16.csv python scripts/profile_forward.py --iters 1 > $O/${S}_ncu_run.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_gemm|mlp_fused|swin_attn" -c 24 \
  -o $O/${S}_gemm_full --force-overwrite python scripts/profile_forward.py --iters 1 > $O/${S}_ncu_full_run.log 2>&1
ncu -i $O/${S}_gemm_full.ncu-rep --page raw --csv > /tmp/${S}_gemm_full_raw.csv 2>/dev/null
python scripts/ncu_gemm_summary.py /tmp/${S}_gemm_full_raw.csv > $O/${S}_gemm_kernels_ncu_full_summary.csv
rm -f $O/${S}_gemm_full.ncu-rep            # (tens of MB with --import-source: gpurun merges at most 64 MiB back)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swin_attn_tc -c 1 -o /tmp/${S}_swin_tc --force-overwrite python scripts/swin_tc_diag.py ncu > $O/${S}_ncu_swin.log 2>&1
python scripts/ncu_summarize.py /tmp/${S}_swin_tc.ncu-rep > $O/${S}_swin_tc_ncu_summary.txt 2>/dev/null
du -sh $O
tail -5 $O/${S}_pytest.log; tail -3 $O/${S}_smoke.log; head -c 1500 $O/${S}_bench_b16.log; echo; head -c 600 $O/${S}_bench_reference_arm.log; echo
echo done > $O/${S}_done.txt
