#!/bin/bash
# round 2, GPU call A: parity of the depth-5 registration kernel + its first A/B + one ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_cloud or align_pose or align_edge or sharded_align_single" > gpurun_out/r2a_pytest_align.log 2>&1
echo "pytest(align subset) rc=$?" | tee -a gpurun_out/r2a_pytest_align.log
CELLS="align_depth=4;align_depth=5;align_depth=5,align_ll=0;align_depth=5,align_resident=0;align_depth=5,align_resident=6;align_depth=5,align_block=512;align_depth=5,l2_persist=0;align_depth=5,align_dynamic_tiles=0"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --shards 8 --cells "$CELLS" --out gpurun_out/r2a_ab_01.json > gpurun_out/r2a_ab_01.log 2>&1
echo "ab 0.1 rc=$?"; tail -20 gpurun_out/r2a_ab_01.log | cut -c1-330
timeout 600 python scripts/ab_r2.py --voxels 0.5 --compact 1 --cells "align_depth=4;align_depth=5;align_depth=5,align_block=512" --out gpurun_out/r2a_ab_05.json > gpurun_out/r2a_ab_05.log 2>&1
echo "ab 0.5 rc=$?"; tail -5 gpurun_out/r2a_ab_05.log | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_align_r2a python scripts/dense_align.py --reps 1 --warmup 2 > gpurun_out/r2a_prof_align.log 2>&1
echo "ncu rc=$?"
timeout 300 python scripts/dense_align.py --sort 1 > gpurun_out/r2a_sort1.log 2>&1; tail -1 gpurun_out/r2a_sort1.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest(all gpu) rc=$?" | tee -a gpurun_out/r2a_pytest_gpu.log
tail -3 gpurun_out/r2a_pytest_gpu.log
