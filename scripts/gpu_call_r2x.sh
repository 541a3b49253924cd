#!/bin/bash
# round 2, GPU call X: phase stamps of the one-cluster voxelize kernel
mkdir -p gpurun_out
for c in 1 8; do
  echo "== ESKF_VOX_CLUSTER=$c"
  ESKF_VOX_CLUSTER=$c ESKF_TRACE=1 timeout 300 python scripts/frame_probe.py 44 2>&1 | grep "voxelize cluster\|preprocess:" | tail -6
done
