#!/bin/bash
# round 2, GPU session 8: VQ-GAN bookends with / without the persistent conv kernel
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s8
RS_CONV_PERSIST=0 timeout 600 python scripts/profile_vq.py > $O/${S}_vq_per_op_nopersist.log 2>&1
RS_CONV_PERSIST=0 RS_CONV_CG=1 timeout 600 python scripts/profile_vq.py > $O/${S}_vq_per_op_nopersist_cg1.log 2>&1
RS_CONV_PERSIST=0 RS_CONV_CG=2 timeout 600 python scripts/profile_vq.py > $O/${S}_vq_per_op_nopersist_cg2.log 2>&1
echo done > $O/${S}_done.txt
