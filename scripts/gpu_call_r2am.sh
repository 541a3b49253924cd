#!/bin/bash
# round 2, GPU call AM: autotune with four candidates: tests + what it picks per size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded_p2p.py -m gpu -x -q -k "large_cloud or sharded or p2p" > gpurun_out/r2am_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2am_pytest.log
ESKF_TRACE=0 timeout 600 python scripts/ab_r2.py --voxels 0.1 --shards 2,4,8 --reps 5 --cells "align_autotune=0;align_autotune=1;align_block=769;align_autotune=1" > gpurun_out/r2am_ab.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2am_ab.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['cell'], d['us_per_iter'], d.get('shard2_us'), d.get('shard4_us'), d.get('shard8_us'), d['ncorr_equal'])
    elif 'PARITY' in l: print(l.strip())
PY
