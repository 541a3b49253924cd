#!/bin/bash
# round 2, GPU call U: stamps_sorted hint (tests), frame breakdown with and without ESKF_TRACE
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_odometry.py -m gpu -x -q -k "stamps or odometry or sequence or crop" > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
timeout 300 python scripts/frame_probe.py 60 > gpurun_out/r2u_probe.log 2>&1; tail -22 gpurun_out/r2u_probe.log
ESKF_TRACE=1 timeout 300 python scripts/frame_probe.py 48 > gpurun_out/r2u_trace.log 2>&1; tail -40 gpurun_out/r2u_trace.log | cut -c1-400
