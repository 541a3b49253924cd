#!/bin/bash
# round 2, GPU call AK: the 3-stage loop on the 8-bit probe filter (align_block 257) as a third autotune candidate
mkdir -p gpurun_out
nvidia-smi --query-gpu=serial --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded_p2p.py tests/test_gpu_baseline_shapes.py -m gpu -x -q -k "large_cloud or sharded or p2p or dense or baseline" > gpurun_out/r2ak_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ak_pytest.log
timeout 900 python scripts/ab_r2.py --voxels 0.1,0.5 --compact 0,1 --shards 8 --reps 5 --cells "align_block=512;align_block=256;align_block=257;align_autotune=1;align_block=512;align_block=257" --out gpurun_out/r2ak_ab.json > gpurun_out/r2ak_ab.log 2>&1
echo "ab rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2ak_ab.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['voxel'], d['compact'], d['cell'], d['us_per_iter'], d.get('shard8_us'), d['ncorr_equal'], d.get('shard8_ok'))
    elif 'PARITY' in l: print(l.strip())
PY
