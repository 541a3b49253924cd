#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_cloud" > gpurun_out/r2t_pytest_align.log 2>&1
echo "pytest(large cloud) rc=$?"; tail -4 gpurun_out/r2t_pytest_align.log
CELLS="align_depth=4;align_depth=8,align_block=512;align_depth=11,align_block=512;align_depth=11,align_block=384;align_depth=12,align_block=512;align_depth=12,align_block=384;!align_depth=11,align_block=384,align_flags=272;!align_depth=11,align_block=384,align_flags=528"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --shards 8 --cells "$CELLS" --out gpurun_out/r2t_ab_01.json 2>&1 | grep "us_per\|PARITY" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'], r.get('shard8_us'), r.get('shard8_ok'))"
timeout 600 python scripts/ab_r2.py --voxels 0.5 --compact 1 --cells "align_depth=4;align_depth=11,align_block=384;align_depth=12,align_block=384" 2>&1 | grep "us_per\|PARITY" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['voxel'], r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'])"
