#!/bin/bash
# round 2, GPU call AL: the 3-stage loop's CTA shapes and ticket settings on whatever kind of box answers
mkdir -p gpurun_out
nvidia-smi --query-gpu=serial --format=csv,noheader
timeout 900 python scripts/ab_r2.py --voxels 0.1 --shards 8 --reps 5 --cells "align_block=512;align_block=256;align_block=384,align_depth=3;align_block=768,align_depth=3;align_block=256,align_ticket_chunk=1;align_block=256,align_ticket_chunk=4;align_block=256,align_dyn16=2;align_block=256,align_dyn16=5;align_block=256,align_dynamic_tiles=0;align_block=256,l2_persist=0;align_block=256" --out gpurun_out/r2al_ab.json > gpurun_out/r2al_ab.log 2>&1
echo "ab rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2al_ab.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['cell'], d['us_per_iter'], d.get('shard8_us'), d['ncorr_equal'])
    elif 'PARITY' in l: print(l.strip())
PY
