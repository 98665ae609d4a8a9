#!/bin/bash
# round 2, GPU session 9: 256x256 VQ-GAN conv layers — timelines, statistics variants, persistent on/off
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s9
TAG=auto timeout 300 python scripts/vq_conv_probe.py > $O/${S}_probe_auto.log 2>&1
TAG=nopersist RS_CONV_PERSIST=0 timeout 300 python scripts/vq_conv_probe.py > $O/${S}_probe_nopersist.log 2>&1
TAG=persist_cg1 RS_CONV_PERSIST=1 RS_CONV_CG=1 timeout 300 python scripts/vq_conv_probe.py > $O/${S}_probe_persist_cg1.log 2>&1
TAG=persist_cg2 RS_CONV_PERSIST=1 RS_CONV_CG=2 timeout 300 python scripts/vq_conv_probe.py > $O/${S}_probe_persist_cg2.log 2>&1
TAG=nopersist_occ1 RS_CONV_PERSIST=0 RS_CONV_OCC=1 timeout 300 python scripts/vq_conv_probe.py > $O/${S}_probe_nopersist_occ1.log 2>&1
echo done > $O/${S}_done.txt
