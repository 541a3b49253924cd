#!/bin/bash
# round 2, GPU call F: the register-resident solve: parity + hand-off stamps + frame bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1
echo "pytest(all gpu) rc=$?"; tail -5 gpurun_out/r2f_pytest_gpu.log
for n in 2000000 250000 15000; do
  echo "== stamps src $n"
  ESKF_ALIGN_STAMPS=1 timeout 300 python scripts/dense_align.py --src $n --reps 1 --warmup 2 2> gpurun_out/r2f_stamps_$n.txt | cut -c1-200
  tail -4 gpurun_out/r2f_stamps_$n.txt
done
CELLS="align_filter=0;align_filter=1,align_flags=16;align_depth=5;align_depth=6;align_depth=7"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0 --shards 8 --cells "$CELLS" --out gpurun_out/r2f_ab_01.json > gpurun_out/r2f_ab_01.log 2>&1
echo "ab 0.1 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2f_ab_01.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'], r.get('shard8_us'), r.get('shard8_ok'))"
