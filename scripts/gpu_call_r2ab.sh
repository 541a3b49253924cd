#!/bin/bash
# round 2, GPU call AB: one-cluster voxelize kernel, trimmed: parity, stamps, frame times (cluster 16 / 8 / grid-wide)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_odometry.py tests/test_golden.py -m gpu -x -q -k "cluster or downsample or preprocess or odometry or sequence or insert or crop or golden or stamps" > gpurun_out/r2ab_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2ab_pytest.log
for c in 1 0 1 0; do
  echo "== ESKF_VOX_CLUSTER=$c"
  ESKF_VOX_CLUSTER=$c ESKF_TRACE=1 timeout 300 python scripts/frame_probe.py 44 2>&1 | grep "voxelize cluster\|preprocess:" | tail -2
  ESKF_VOX_CLUSTER=$c timeout 300 python scripts/frame_probe.py 60 | tail -20 | awk '{s+=$5; n++} END {printf "mean dev ms over last %d frames: %.4f\n", n, s/n}'
done
