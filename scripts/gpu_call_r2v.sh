#!/bin/bash
# round 2, GPU call V: one-cluster voxelize kernel: parity vs the grid-wide kernel, frame times per setting
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_odometry.py -m gpu -x -q -k "cluster or downsample or preprocess or odometry or sequence or insert or crop" > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2v_pytest.log
for c in 0 1 8; do
  echo "== ESKF_VOX_CLUSTER=$c"
  ESKF_VOX_CLUSTER=$c timeout 300 python scripts/frame_probe.py 56 > gpurun_out/r2v_probe_$c.log 2>&1; tail -6 gpurun_out/r2v_probe_$c.log
  ESKF_VOX_CLUSTER=$c ESKF_TRACE=1 timeout 300 python scripts/frame_probe.py 44 2>&1 | grep "preprocess:" | tail -3
done
