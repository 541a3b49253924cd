#!/bin/bash
# round 2, GPU call AC: flagged-word pose broadcast for the default kernels (depth 3 / 4) + level tables sized by kept points
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ac_pytest.log
timeout 600 python scripts/ab_r2.py --voxels 0.1 --shards 8 --cells "align_ll=0;align_ll=1;align_ll=0;align_ll=1" --out gpurun_out/r2ac_ab.json > gpurun_out/r2ac_ab.log 2>&1
echo "ab rc=$?"; cut -c1-260 gpurun_out/r2ac_ab.log | tail -6
for v in 0 1 0 1; do
  echo "== ESKF_ALIGN_LL=$v"
  ESKF_ALIGN_LL=$v timeout 300 python scripts/frame_probe.py 60 | tail -20 | awk '{s+=$5; n++; f+=$8} END {printf "mean dev ms over last %d frames: %.4f  filter stage %.4f\n", n, s/n, f/n}'
done
ESKF_TRACE=1 timeout 300 python scripts/frame_probe.py 44 2>&1 | grep "preprocess:\|align:\|map_insert:" | tail -6
