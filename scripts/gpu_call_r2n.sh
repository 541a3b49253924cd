#!/bin/bash
# depth 10 (three launches per iteration): parity, A/B, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_cloud or align_pose or align_edge or sharded_align_single" > gpurun_out/r2n_pytest_align.log 2>&1
echo "pytest(align subset) rc=$?"; tail -5 gpurun_out/r2n_pytest_align.log
CELLS="align_depth=4;align_depth=8,align_block=512;align_depth=10;align_depth=10,align_flags=20;align_depth=10,align_flags=21"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --shards 8 --cells "$CELLS" --out gpurun_out/r2n_ab_01.json > gpurun_out/r2n_ab_01.log 2>&1
echo "ab 0.1 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2n_ab_01.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'], r.get('shard8_us'), r.get('shard8_ok'))"
timeout 600 python scripts/ab_r2.py --voxels 0.5 --compact 1 --cells "align_depth=4;align_depth=10" --out gpurun_out/r2n_ab_05.json > gpurun_out/r2n_ab_05.log 2>&1
echo "ab 0.5 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2n_ab_05.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'])"
ESKF_ALIGN_DEPTH=10 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mk_ -s 60 -c 30 --csv --log-file gpurun_out/r2n_ncu.csv python scripts/dense_align.py --reps 1 --warmup 2 > gpurun_out/r2n_ncu_run.log 2>&1
echo "ncu rc=$?"
