#!/bin/bash
# round 2, GPU call AE: 7-neighbour registration, serial vs batched key loads (+ sector prefetch), same process
mkdir -p gpurun_out
timeout 1500 python scripts/ab_nn7.py --voxels 0.1,0.5 > gpurun_out/r2ae_nn7.log 2>&1
echo "rc=$?"; cut -c1-200 gpurun_out/r2ae_nn7.log
