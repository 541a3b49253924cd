#!/bin/bash
# round 2, GPU call E: depth 4 on the 8-bit filter; where does an iteration's hand-off time go
mkdir -p gpurun_out
CELLS="align_filter=0;align_filter=1;align_filter=1,align_flags=16;align_filter=1,align_flags=16,l2_persist=0;align_filter=1,align_flags=20;align_filter=1,align_flags=21;align_block=512,align_filter=1,align_flags=16;align_filter=0,l2_persist=0"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --shards 8 --cells "$CELLS" --out gpurun_out/r2e_ab_01.json > gpurun_out/r2e_ab_01.log 2>&1
echo "ab 0.1 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2e_ab_01.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'], r.get('shard8_us'), r.get('shard8_ok'))"
timeout 600 python scripts/ab_r2.py --voxels 0.5 --compact 1 --cells "align_filter=0;align_filter=1;align_filter=1,align_flags=16" --out gpurun_out/r2e_ab_05.json > gpurun_out/r2e_ab_05.log 2>&1
echo "ab 0.5 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2e_ab_05.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'])"
for n in 2000000 250000 65536; do
  echo "== stamps src $n"
  ESKF_ALIGN_STAMPS=1 timeout 300 python scripts/dense_align.py --src $n --reps 1 --warmup 2 2> gpurun_out/r2e_stamps_$n.txt | cut -c1-200
  tail -10 gpurun_out/r2e_stamps_$n.txt
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:align_kernel --csv --log-file gpurun_out/r2e_ncu.csv \
    python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --warm 1 --reps 1 --cells "$CELLS" > gpurun_out/r2e_ncu_run.log 2>&1
echo "ncu rc=$?"
