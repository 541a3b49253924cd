#!/bin/bash
mkdir -p gpurun_out
CELLS="align_depth=4;align_dyn16=2;align_dyn16=4;align_dyn16=5;align_dyn16=6;align_dyn16=8;align_dyn16=4,align_ticket_chunk=1;align_dyn16=6,align_ticket_chunk=1;align_dyn16=6,align_ticket_chunk=4"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0 --shards 8 --cells "$CELLS" 2>&1 | grep "us_per\|PARITY" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'], r.get('shard8_us'), r.get('shard8_ok'))"
