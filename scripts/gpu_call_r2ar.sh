#!/bin/bash
# round 2, GPU call AR: L2 position prefetch in the 3-stage loop of the fat CTAs (align_flags bit 2048 = off)
mkdir -p gpurun_out
nvidia-smi --query-gpu=serial --format=csv,noheader
timeout 600 python scripts/ab_r2.py --voxels 0.1 --shards 8 --reps 7 --cells "align_block=769;align_block=769,align_flags=2064;align_block=769;align_block=769,align_flags=2064;align_block=512;align_autotune=1" > gpurun_out/r2ar_ab.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2ar_ab.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['cell'], d['us_per_iter'], d.get('shard8_us'), d['ncorr_equal'])
    elif 'PARITY' in l: print(l.strip())
PY
