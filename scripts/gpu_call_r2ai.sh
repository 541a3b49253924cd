#!/bin/bash
# round 2, GPU call AI: large-cloud autotune: tests, what it picks here, and the numbers either way
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ai_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ai_pytest.log
timeout 600 python scripts/ab_r2.py --voxels 0.1 --shards 8 --reps 5 --cells "align_autotune=0;align_autotune=1;align_block=256;align_autotune=0;align_autotune=1" --out gpurun_out/r2ai_ab.json > gpurun_out/r2ai_ab.log 2>&1
echo "ab rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2ai_ab.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['cell'], d['us_per_iter'], d.get('shard8_us'), d['ncorr_equal'])
PY
