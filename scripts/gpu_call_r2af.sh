#!/bin/bash
# round 2, GPU call AF: which kernel shape for a 250 k-point shard (8-way split of the dense source) and a 500 k / 1 M one
mkdir -p gpurun_out
timeout 900 python scripts/ab_r2.py --voxels 0.1 --shards 2,4,8 --reps 7 --cells "align_fat_points=131072;align_fat_points=300000;align_fat_points=600000;align_fat_points=1200000;align_fat_points=131072,align_block=384;align_fat_points=131072,align_dynamic_tiles=0" --out gpurun_out/r2af_ab.json > gpurun_out/r2af_ab.log 2>&1
echo "ab rc=$?"; cut -c1-330 gpurun_out/r2af_ab.log | tail -8
