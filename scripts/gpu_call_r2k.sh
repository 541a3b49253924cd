#!/bin/bash
# round 2, GPU call K: the depth-8 (parked candidates) registration kernel: parity, A/B, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_cloud or align_pose or align_edge or sharded_align_single" > gpurun_out/r2k_pytest_align.log 2>&1
echo "pytest(align subset) rc=$?"; tail -5 gpurun_out/r2k_pytest_align.log
CELLS="align_depth=4;align_depth=8;align_depth=8,align_block=512;align_depth=8,align_block=768;align_depth=8,align_dynamic_tiles=0;align_depth=8,align_flags=20"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --shards 8 --cells "$CELLS" --out gpurun_out/r2k_ab_01.json > gpurun_out/r2k_ab_01.log 2>&1
echo "ab 0.1 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2k_ab_01.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'], r.get('shard8_us'), r.get('shard8_ok'))"
timeout 600 python scripts/ab_r2.py --voxels 0.5 --compact 1 --cells "align_depth=4;align_depth=8;align_depth=8,align_block=512;align_depth=8,align_block=768" --out gpurun_out/r2k_ab_05.json > gpurun_out/r2k_ab_05.log 2>&1
echo "ab 0.5 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2k_ab_05.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'])"
ESKF_ALIGN_DEPTH=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_align_r2k python scripts/dense_align.py --reps 1 --warmup 2 > gpurun_out/r2k_prof_align.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/r2k_prof_align.log | cut -c1-300
