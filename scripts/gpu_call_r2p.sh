#!/bin/bash
mkdir -p gpurun_out
CELLS="align_depth=4;align_depth=4,align_block=512;align_depth=8,align_block=512;align_depth=8,align_block=640;align_depth=4;align_depth=8,align_block=512"
timeout 600 python scripts/ab_r2.py --voxels 0.1 --compact 0,1 --shards 8 --cells "$CELLS" --out gpurun_out/r2p_ab_01.json > gpurun_out/r2p_ab_01.log 2>&1
echo "ab 0.1 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2p_ab_01.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'], r.get('shard8_us'), r.get('shard8_ok'))"
timeout 600 python scripts/ab_r2.py --voxels 0.5,1.0 --compact 1 --cells "align_depth=4;align_depth=8,align_block=512" --out gpurun_out/r2p_ab_05.json > gpurun_out/r2p_ab_05.log 2>&1
echo "ab 0.5/1.0 rc=$?"; grep "us_per\|PARITY" gpurun_out/r2p_ab_05.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('PARITY'): print(l.strip()); continue
    r = json.loads(l); print(r['voxel'], r['compact'], r['cell'], r['us_per_iter'], r['ncorr_equal'])"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest_gpu.log 2>&1
echo "pytest(all gpu) rc=$?"; tail -4 gpurun_out/r2p_pytest_gpu.log
