#!/bin/bash
# GPU session 19: ncu --set full on the fused MLP, GroupNorm apply and window attention kernels (one forward, batch 16)
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mlp_fused -c 3 \
    -o gpurun_out/prof_mlp -f python scripts/profile_forward.py --iters 1 > gpurun_out/ncu_mlp.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_apply -c 12 \
    -o gpurun_out/prof_gn -f python scripts/profile_forward.py --iters 1 > gpurun_out/ncu_gn.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:window_attn -c 3 \
    -o gpurun_out/prof_attn -f python scripts/profile_forward.py --iters 1 > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_mlp.log gpurun_out/ncu_gn.log gpurun_out/ncu_attn.log
ls -la gpurun_out/*.ncu-rep
