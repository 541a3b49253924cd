#!/bin/bash
# round 2, GPU call AN: the ncu capture of the roofline kernel again (the launch AFTER the autotune and the warm-ups)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 18 -c 1 -f \
    -o gpurun_out/prof_align_r2 python scripts/dense_align.py --reps 1 --warmup 2 > gpurun_out/prof_align_r2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/prof_align_r2.log | cut -c1-300
nvidia-smi --query-gpu=serial --format=csv,noheader
