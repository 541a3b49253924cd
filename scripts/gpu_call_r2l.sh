#!/bin/bash
# ablation of the depth-8 kernel (512 threads): where does a pass go?
mkdir -p gpurun_out
B="align_depth=8,align_block=512"
CELLS="align_depth=4;$B;!$B,align_flags=2064;!$B,align_flags=272;!$B,align_flags=1296;!$B,align_flags=3344;!$B,align_flags=528;!$B,align_flags=1552"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0 --shards 8 --cells "$CELLS" --out gpurun_out/r2l_ab_01.json > gpurun_out/r2l_ab_01.log 2>&1
echo "ab 0.1 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2l_ab_01.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r.get('shard8_us'))"
for f in 16 272 1296; do
  echo "== stamps flags $f"
  ESKF_ALIGN_DEPTH=8 ESKF_ALIGN_BLOCK=512 ESKF_ALIGN_FLAGS=$f ESKF_ALIGN_STAMPS=1 timeout 300 python scripts/dense_align.py --reps 1 --warmup 2 2> gpurun_out/r2l_stamps_$f.txt > /dev/null
  tail -3 gpurun_out/r2l_stamps_$f.txt
done
