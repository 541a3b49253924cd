#!/bin/bash
# round 2, GPU call AP: k-NN level tables built by the one-cluster voxelize kernel itself: parity + frame times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_odometry.py tests/test_golden.py -m gpu -x -q -k "cluster or downsample or preprocess or odometry or sequence or crop or golden or stamps" > gpurun_out/r2ap_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2ap_pytest.log
ESKF_TRACE=1 timeout 300 python scripts/frame_probe.py 44 2>&1 | grep "voxelize cluster\|preprocess:" | tail -2
timeout 300 python scripts/frame_probe.py 60 | tail -20 | awk '{s+=$5; n++} END {printf "mean dev ms over last %d frames: %.4f\n", n, s/n}'
ESKF_VOX_CLUSTER=0 timeout 300 python scripts/frame_probe.py 60 | tail -20 | awk '{s+=$5; n++} END {printf "grid-wide kernel: mean dev ms over last %d frames: %.4f\n", n, s/n}'
