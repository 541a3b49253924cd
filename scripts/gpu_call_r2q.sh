#!/bin/bash
mkdir -p gpurun_out
V=$PWD/eskf_lio_b200/lib/variants
CELLS="align_depth=4;align_depth=4,align_block=448;align_depth=4,align_block=384;align_depth=4,align_block=640;align_depth=4,align_ticket_chunk=1;align_depth=4,align_ticket_chunk=4;align_depth=4,align_flags=20"
for v in default posahead pospf4 posahead_pf; do
  if [ $v = default ]; then lib=""; else lib=$V/$v.so; fi
  echo "== $v"
  ESKF_GPU_LIB=$lib timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0 --shards 8 --cells "$CELLS" 2>&1 | grep "us_per\|PARITY" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'], r.get('shard8_us'), r.get('shard8_ok'))"
done
