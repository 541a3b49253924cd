#!/bin/bash
# round 2, GPU call B: what does the L2 keep between Gauss-Newton iterations?  (cache-policy cells, timed and under ncu)
mkdir -p gpurun_out
CELLS="align_depth=4;align_depth=5;align_flags=4;align_flags=5;align_flags=20;align_flags=21;align_flags=53;align_flags=53,l2_persist=0;align_flags=85;align_flags=16;l2_persist=0;align_flags=117;align_flags=117,align_resident=0"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --cells "$CELLS" --out gpurun_out/r2b_ab.json > gpurun_out/r2b_ab.log 2>&1
echo "ab rc=$?"; grep us_per gpurun_out/r2b_ab.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'])"
ESKF_L2_CARVEOUT=0 timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --cells "align_depth=4;align_depth=5;align_flags=5;align_flags=21;align_flags=53;align_flags=117" --out gpurun_out/r2b_ab_nocarve.json > gpurun_out/r2b_ab_nocarve.log 2>&1
echo "ab nocarve rc=$?"; grep us_per gpurun_out/r2b_ab_nocarve.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('nocarve', r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'])"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:align_kernel --csv --log-file gpurun_out/r2b_ncu.csv \
    python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --warm 1 --reps 1 --cells "$CELLS" > gpurun_out/r2b_ncu_run.log 2>&1
echo "ncu rc=$?"
ESKF_L2_CARVEOUT=0 timeout 900 ncu --metrics $M --clock-control none -k regex:align_kernel --csv --log-file gpurun_out/r2b_ncu_nocarve.csv \
    python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --warm 1 --reps 1 --cells "align_depth=4;align_depth=5;align_flags=5;align_flags=21;align_flags=53;align_flags=117" > gpurun_out/r2b_ncu_nocarve_run.log 2>&1
echo "ncu nocarve rc=$?"
