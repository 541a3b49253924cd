#!/bin/bash
# round 2, GPU call Z: one-cluster voxelize kernel after the per-phase rework: parity, stamps, frame times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_odometry.py tests/test_golden.py -m gpu -x -q -k "cluster or downsample or preprocess or odometry or sequence or insert or crop or golden or stamps" > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2z_pytest.log
for c in 1 8 0; do
  echo "== ESKF_VOX_CLUSTER=$c"
  ESKF_VOX_CLUSTER=$c ESKF_TRACE=1 timeout 300 python scripts/frame_probe.py 44 2>&1 | grep "voxelize cluster\|preprocess:" | tail -4
  ESKF_VOX_CLUSTER=$c timeout 300 python scripts/frame_probe.py 56 > gpurun_out/r2z_probe_$c.log 2>&1; tail -5 gpurun_out/r2z_probe_$c.log
done
